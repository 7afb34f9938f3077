mkdir -p gpurun_out
run() { # tag flags workload
  GORP_TAIL_FLAGS=$2 python bench.py --workload $3 --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" --lines-per-gpu 40000000 > gpurun_out/r2q_bench_$3_$1.json 2>> gpurun_out/r2q_err.txt
}
for W in syslog200 weblog; do
  run base 0 $W
  run sortg8 $((16 + 8*256)) $W
  run sortg10 $((16 + 10*256)) $W
  run sortg12 $((16 + 12*256)) $W
done
python -m pytest tests -m gpu -x -q -k "config_corpora or tricky or long or edge" > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
GORP_TAIL_FLAGS=$((16 + 8*256)) python -m pytest tests -m gpu -x -q -k "config_corpora or long" > gpurun_out/r2q_pytest_sorted.log 2>&1; tail -3 gpurun_out/r2q_pytest_sorted.log
tail -c 400 gpurun_out/r2q_err.txt
