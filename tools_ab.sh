mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "k1h or config_corpora or edge or long or non_ascii" > gpurun_out/r3w_pytest.log 2>&1; tail -n 6 gpurun_out/r3w_pytest.log
for W in syslog200 weblog utf16mix; do
python bench.py --workload $W --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" --lines-per-gpu 40000000 > gpurun_out/r3w_bench_$W.json 2>> gpurun_out/r3w_err.txt
GORP_NO_HEADWALK=1 python bench.py --workload $W --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" --lines-per-gpu 40000000 > gpurun_out/r3w_bench_${W}_nohw.json 2>> gpurun_out/r3w_err.txt
done
tail -c 400 gpurun_out/r3w_err.txt
