mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r3u_pytest.log 2>&1; tail -n 4 gpurun_out/r3u_pytest.log
python bench.py > gpurun_out/r3u_bench_n1.json 2> gpurun_out/r3u_bench_n1.err; tail -c 300 gpurun_out/r3u_bench_n1.err
python bench.py --impl reference > gpurun_out/r3u_bench_ref.json 2> gpurun_out/r3u_bench_ref.err; tail -c 300 gpurun_out/r3u_bench_ref.err
W=syslog200
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3u_launches_$W.csv \
   python bench.py --workload $W --steps 1 --warmup 3 --lines-per-gpu 16000000 --skip-e2e --skip-cpu --configs "" > gpurun_out/r3u_launches_$W.json 2> gpurun_out/r3u_launches_$W.err
ncu --set full --clock-control none --import-source on -k regex:"tailwalk_kernel" -s 3 -c 1 -f -o gpurun_out/r3u_prof_$W \
   python bench.py --workload $W --steps 1 --warmup 3 --lines-per-gpu 16000000 --skip-e2e --skip-cpu --configs "" > gpurun_out/r3u_prof_$W.log 2>&1
