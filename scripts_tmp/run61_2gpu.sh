timeout 600 python profiles/r2_cfg5.py 1e8 2 > gpurun_out/r2f_cfg5_2gpu.json 2> gpurun_out/r2f_cfg5_2gpu.err; cat gpurun_out/r2f_cfg5_2gpu.json
COAL_FUZZ_N=300 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "random_problems" > gpurun_out/r2f_fuzz300.log 2>&1; echo "pytest rc=$? (COAL_FUZZ_N=300)" >> gpurun_out/r2f_fuzz300.log; tail -4 gpurun_out/r2f_fuzz300.log
