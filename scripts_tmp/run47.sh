M="--metrics gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-e2e"
timeout 600 ncu $M --clock-control none -k regex:history_lockstep -s 48 -c 24 --csv --log-file gpurun_out/r2_gcc_def.csv $B > /dev/null 2>&1
COALGPU_LOCKSTEP_NH=32 COALGPU_LOCKSTEP_NT=128 timeout 600 ncu $M --clock-control none -k regex:history_lockstep -s 48 -c 24 --csv --log-file gpurun_out/r2_gcc_32128.csv $B > /dev/null 2>&1
COALGPU_LIB=/root/repo/variants/u11.so timeout 600 ncu $M --clock-control none -k regex:history_lockstep -s 48 -c 24 --csv --log-file gpurun_out/r2_gcc_u11.csv $B > /dev/null 2>&1
COALGPU_LIB=/root/repo/variants/u44.so timeout 600 ncu $M --clock-control none -k regex:history_lockstep -s 48 -c 24 --csv --log-file gpurun_out/r2_gcc_u44.csv $B > /dev/null 2>&1
ls -la gpurun_out/r2_gcc_*.csv
