set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_p_tests.log 2>&1; tail -5 gpurun_out/r2_p_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
for rep in 1 2; do
$B --workload cfg3 > gpurun_out/r2_p_cfg3_$rep.log 2>&1
SSR_B200_CTL_UPLOAD=copy $B --workload cfg3 > gpurun_out/r2_p_cfg3_copy_$rep.log 2>&1
$B --workload cfg2 > gpurun_out/r2_p_cfg2_$rep.log 2>&1
SSR_B200_CTL_UPLOAD=copy $B --workload cfg2 > gpurun_out/r2_p_cfg2_copy_$rep.log 2>&1
done
SSR_B200_DYN_TMA=0 $B --workload cfg3 > gpurun_out/r2_p_cfg3_oldkernels.log 2>&1
SSR_B200_DYN_TMA=0 $B --workload cfg2 > gpurun_out/r2_p_cfg2_oldkernels.log 2>&1
python scripts/summarize.py gpurun_out/r2_p_cfg*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_p_*.log | head
