mkdir -p gpurun_out
for T in 512 384 256; do
  GORP_CW_THREADS=$T python bench.py --workload readme --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" > gpurun_out/r2z_bench_readme_t$T.json 2>> gpurun_out/r2z_err.txt
  GORP_CW_THREADS=$T python bench.py --workload simple --lines-per-gpu 40000000 --steps 10 --warmup 3 --skip-e2e --skip-cpu --configs "" > gpurun_out/r2z_bench_simple_t$T.json 2>> gpurun_out/r2z_err.txt
done
GORP_CW_THREADS=384 GORP_ONEPASS_DEBUG=1 python bench.py --workload readme --steps 3 --warmup 2 --skip-e2e --skip-cpu --configs "" 2>&1 >/dev/null | grep "thread-cycles" | tail -1
GORP_CW_THREADS=256 GORP_ONEPASS_DEBUG=1 python bench.py --workload readme --steps 3 --warmup 2 --skip-e2e --skip-cpu --configs "" 2>&1 >/dev/null | grep "thread-cycles" | tail -1
tail -c 300 gpurun_out/r2z_err.txt
