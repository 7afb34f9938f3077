mkdir -p gpurun_out
run() { # tag flags workload
  GORP_TAIL_FLAGS=$2 python bench.py --workload $3 --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" --lines-per-gpu 40000000 > gpurun_out/r2x_bench_$3_$1.json 2>> gpurun_out/r2x_err.txt
}
for W in syslog200 weblog utf16mix; do
  run new 0 $W
  run oldstore 512 $W
done
python -m pytest tests -m gpu -x -q -k "config_corpora or tricky or long or edge" > gpurun_out/r2x_pytest.log 2>&1; tail -n 3 gpurun_out/r2x_pytest.log
tail -c 400 gpurun_out/r2x_err.txt
