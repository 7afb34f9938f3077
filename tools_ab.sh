mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r3n_pytest.log 2>&1; tail -n 5 gpurun_out/r3n_pytest.log
python bench.py --steps 5 --warmup 3 --skip-cpu --configs "" > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_err.txt
tail -c 300 gpurun_out/r3n_err.txt
