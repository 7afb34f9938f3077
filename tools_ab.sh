mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tricky or long or edge or (tier and (fusedwalk or tailwalk or chunkwalk)) or config_corpora or non_ascii" > gpurun_out/r3e_pytest.log 2>&1; tail -n 4 gpurun_out/r3e_pytest.log
GORP_SMALL_PATH=fusedwalk python bench.py --workload readme --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" > gpurun_out/r3e_bench_readme_fusedwalk.json 2>> gpurun_out/r3e_err.txt
GORP_SMALL_PATH=fusedwalk python bench.py --workload simple --lines-per-gpu 40000000 --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" > gpurun_out/r3e_bench_simple_fusedwalk.json 2>> gpurun_out/r3e_err.txt
for W in syslog200 weblog utf16mix; do
python bench.py --workload $W --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" --lines-per-gpu 40000000 > gpurun_out/r3e_bench_$W.json 2>> gpurun_out/r3e_err.txt
done
tail -c 600 gpurun_out/r3e_err.txt
