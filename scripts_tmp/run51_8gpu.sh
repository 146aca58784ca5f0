nvidia-smi -L | wc -l > gpurun_out/r2f_8gpu_devices.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2f_bench_8gpu.json 2> gpurun_out/r2f_bench_8gpu.err; tail -c 300 gpurun_out/r2f_bench_8gpu.json
timeout 600 python profiles/r2_cfg5.py 1e8 8 > gpurun_out/r2f_cfg5_8gpu.json 2> gpurun_out/r2f_cfg5_8gpu.err; cat gpurun_out/r2f_cfg5_8gpu.json
timeout 600 python profiles/r2_cfg5.py 1e8 4 > gpurun_out/r2f_cfg5_4gpu.json 2> gpurun_out/r2f_cfg5_4gpu.err; cat gpurun_out/r2f_cfg5_4gpu.json
