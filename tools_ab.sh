mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_device or sharding or concurrent" > gpurun_out/r3v_pytest_2gpu.log 2>&1; tail -n 3 gpurun_out/r3v_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r3v_bench_n2.json 2> gpurun_out/r3v_bench_n2.err; tail -c 400 gpurun_out/r3v_bench_n2.err
