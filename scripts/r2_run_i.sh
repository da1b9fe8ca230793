set -x
python scripts/mac_trace.py 64 2>&1 | grep block
python scripts/mac_trace.py 512 2>&1 | grep block
python -m pytest tests -x -q -m gpu > gpurun_out/r2_i_tests.log 2>&1; tail -3 gpurun_out/r2_i_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --outputs 64 > gpurun_out/r2_i_shard64.log 2>&1
SSR_B200_MAC_CALIBRATE=0 $B --outputs 64 > gpurun_out/r2_i_shard64_nocal.log 2>&1
$B --steps 300 > gpurun_out/r2_i_cfg5.log 2>&1
SSR_B200_MAC_CALIBRATE=0 $B --steps 300 > gpurun_out/r2_i_cfg5_nocal.log 2>&1
$B --workload cfg4 > gpurun_out/r2_i_cfg4.log 2>&1
SSR_B200_MAC_CALIBRATE=0 $B --workload cfg4 > gpurun_out/r2_i_cfg4_nocal.log 2>&1
python scripts/summarize.py gpurun_out/r2_i_*.log
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_i_c*.log gpurun_out/r2_i_s*.log | head
