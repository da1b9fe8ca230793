set -x
python -m pytest tests/test_gpu_nonuniform.py tests/test_gpu_renderers.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r2_o_tests.log 2>&1; tail -3 gpurun_out/r2_o_tests.log
SSR_B200_TRACE_ISSUE=1 python scripts/mac_trace2.py cfg3 > gpurun_out/r2_o_trace_cfg3.log 2>&1
tail -9 gpurun_out/r2_o_trace_cfg3.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --workload cfg3 > gpurun_out/r2_o_cfg3.log 2>&1
$B --workload cfg2 > gpurun_out/r2_o_cfg2.log 2>&1
$B --workload cfg4 > gpurun_out/r2_o_cfg4.log 2>&1
python scripts/summarize.py gpurun_out/r2_o_cfg*.log
$B --workload cfg4 --nonuniform 1x4,2x2 > gpurun_out/r2_o_nu_cfg4.log 2>&1; tail -2 gpurun_out/r2_o_nu_cfg4.log | cut -c1-1500
$B --nonuniform 1x8,4x10 --outputs 64 > gpurun_out/r2_o_nu_shard64.log 2>&1; tail -2 gpurun_out/r2_o_nu_shard64.log | cut -c1-1500
$B --nonuniform 1x8,4x10 > gpurun_out/r2_o_nu_cfg5.log 2>&1; tail -2 gpurun_out/r2_o_nu_cfg5.log | cut -c1-1500
$B --nonuniform 1x4,2x2,4x10 > gpurun_out/r2_o_nu3_cfg5.log 2>&1; tail -2 gpurun_out/r2_o_nu3_cfg5.log | cut -c1-1500
