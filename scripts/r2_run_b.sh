set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_b_tests.log 2>&1; tail -3 gpurun_out/r2_b_tests.log
B="python bench.py --no-cpu-baseline --steps 300 --warmup 50"
$B > gpurun_out/r2_b_cfg5.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B > gpurun_out/r2_b_cfg5_nooverlap.log 2>&1
$B --outputs 64 > gpurun_out/r2_b_shard64.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --outputs 64 > gpurun_out/r2_b_shard64_nooverlap.log 2>&1
$B --outputs 64 --no-stage-timing > gpurun_out/r2_b_shard64_nostage.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --outputs 64 --no-stage-timing > gpurun_out/r2_b_shard64_nostage_nooverlap.log 2>&1
$B --workload cfg4 --steps 1000 > gpurun_out/r2_b_cfg4.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg4 --steps 1000 > gpurun_out/r2_b_cfg4_nooverlap.log 2>&1
$B --workload cfg4 --steps 1000 --no-stage-timing > gpurun_out/r2_b_cfg4_nostage.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg4 --steps 1000 --no-stage-timing > gpurun_out/r2_b_cfg4_nostage_nooverlap.log 2>&1
python scripts/summarize.py gpurun_out/r2_b_*.log
