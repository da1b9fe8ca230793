set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_e_tests.log 2>&1; tail -3 gpurun_out/r2_e_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100 --no-stage-timing"
$B --workload cfg3 > gpurun_out/r2_e_cfg3.log 2>&1
SSR_B200_DYN_CTAS=296 $B --workload cfg3 > gpurun_out/r2_e_cfg3_g296.log 2>&1
SSR_B200_DYN_CTAS=264 $B --workload cfg3 > gpurun_out/r2_e_cfg3_g264.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg3 > gpurun_out/r2_e_cfg3_nooverlap.log 2>&1
$B --workload cfg3 > gpurun_out/r2_e_cfg3_again.log 2>&1
$B --workload cfg2 > gpurun_out/r2_e_cfg2.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg2 > gpurun_out/r2_e_cfg2_nooverlap.log 2>&1
python scripts/summarize.py gpurun_out/r2_e_*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_e_cfg*.log | head
