set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_c_tests.log 2>&1; tail -5 gpurun_out/r2_c_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --workload cfg3 > gpurun_out/r2_c_cfg3.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg3 > gpurun_out/r2_c_cfg3_nooverlap.log 2>&1
$B --workload cfg3 --no-stage-timing > gpurun_out/r2_c_cfg3_nostage.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg3 --no-stage-timing > gpurun_out/r2_c_cfg3_nostage_nooverlap.log 2>&1
$B --workload cfg3 --static-head > gpurun_out/r2_c_cfg3_static.log 2>&1
$B --workload cfg2 > gpurun_out/r2_c_cfg2.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg2 > gpurun_out/r2_c_cfg2_nooverlap.log 2>&1
$B --workload cfg2 --no-stage-timing > gpurun_out/r2_c_cfg2_nostage.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg2 --no-stage-timing > gpurun_out/r2_c_cfg2_nostage_nooverlap.log 2>&1
python scripts/summarize.py gpurun_out/r2_c_*.log
grep -h "Error\|error\|Traceback" gpurun_out/r2_c_cfg*.log | head
