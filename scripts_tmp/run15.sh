timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_full7.log; tail -25 gpurun_out/r2_pytest_full7.log
bash scripts_tmp/run14.sh
