set -x
python -m pytest tests/test_gpu_nonuniform.py -x -q -m gpu > gpurun_out/r2_n_nu_tests.log 2>&1; tail -15 gpurun_out/r2_n_nu_tests.log
SSR_B200_TRACE_ISSUE=1 python scripts/mac_trace2.py cfg3 > gpurun_out/r2_n_trace_cfg3.log 2>&1
SSR_B200_TRACE_ISSUE=1 SSR_B200_DYN_CLUSTER=8 python scripts/mac_trace2.py cfg3 > gpurun_out/r2_n_trace_cfg3_c8.log 2>&1
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --workload cfg3 > gpurun_out/r2_n_cfg3.log 2>&1
$B --workload cfg4 > gpurun_out/r2_n_cfg4.log 2>&1
python scripts/summarize.py gpurun_out/r2_n_cfg*.log
tail -12 gpurun_out/r2_n_trace_cfg3.log; tail -9 gpurun_out/r2_n_trace_cfg3_c8.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2_n_tests.log 2>&1; tail -5 gpurun_out/r2_n_tests.log
