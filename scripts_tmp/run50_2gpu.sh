nvidia-smi -L > gpurun_out/r2f_2gpu_devices.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_2gpu.log; tail -6 gpurun_out/r2f_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; tail -c 600 gpurun_out/r2f_bench_2gpu.json
