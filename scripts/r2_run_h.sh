set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_h_tests.log 2>&1; tail -3 gpurun_out/r2_h_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --workload cfg4 > gpurun_out/r2_h_cfg4.log 2>&1
SSR_B200_ZEROCOPY_IN=0 $B --workload cfg4 > gpurun_out/r2_h_cfg4_nozcin.log 2>&1
$B --outputs 64 > gpurun_out/r2_h_shard64.log 2>&1
SSR_B200_ZEROCOPY_IN=0 $B --outputs 64 > gpurun_out/r2_h_shard64_nozcin.log 2>&1
$B --workload cfg3 > gpurun_out/r2_h_cfg3.log 2>&1
SSR_B200_ZEROCOPY_IN=0 $B --workload cfg3 > gpurun_out/r2_h_cfg3_nozcin.log 2>&1
$B --workload cfg2 > gpurun_out/r2_h_cfg2.log 2>&1
SSR_B200_ZEROCOPY_IN=0 $B --workload cfg2 > gpurun_out/r2_h_cfg2_nozcin.log 2>&1
python scripts/summarize.py gpurun_out/r2_h_*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_h_c*.log gpurun_out/r2_h_s*.log | head
bash profiles/ncu_capture_r2.sh 2>&1 | tail -45
