set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_l_tests.log 2>&1; tail -5 gpurun_out/r2_l_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --workload cfg3 > gpurun_out/r2_l_cfg3.log 2>&1
SSR_B200_DYN_TMA=0 $B --workload cfg3 > gpurun_out/r2_l_cfg3_old.log 2>&1
SSR_B200_DYN_CLUSTER=8 $B --workload cfg3 > gpurun_out/r2_l_cfg3_c8.log 2>&1
$B --workload cfg2 > gpurun_out/r2_l_cfg2.log 2>&1
SSR_B200_DYN_TMA=0 $B --workload cfg2 > gpurun_out/r2_l_cfg2_old.log 2>&1
$B --workload cfg4 > gpurun_out/r2_l_cfg4.log 2>&1
$B --workload cfg4 --no-stage-timing > gpurun_out/r2_l_cfg4_nostage.log 2>&1
$B --outputs 64 > gpurun_out/r2_l_shard64.log 2>&1
$B --outputs 64 --no-stage-timing > gpurun_out/r2_l_shard64_nostage.log 2>&1
python scripts/summarize.py gpurun_out/r2_l_*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_l_*.log | head
