set -x
python -m pytest tests -x -q -m gpu -k "renderers or golden or dropin" > gpurun_out/r2_m_tests.log 2>&1; tail -3 gpurun_out/r2_m_tests.log
SSR_B200_TRACE_ISSUE=1 python scripts/mac_trace2.py cfg4 > gpurun_out/r2_m_trace_cfg4.log 2>&1
SSR_B200_TRACE_ISSUE=1 python scripts/mac_trace2.py cfg3 > gpurun_out/r2_m_trace_cfg3.log 2>&1
SSR_B200_TRACE_ISSUE=1 python scripts/mac_trace2.py cfg5 64 > gpurun_out/r2_m_trace_shard64.log 2>&1
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
OLD=$PWD/ssr-pre-0.4.0_b200/lib/libssr_b200_head.so
for rep in 1 2; do
$B --workload cfg3 > gpurun_out/r2_m_cfg3_$rep.log 2>&1
SSR_B200_LIB=$OLD $B --workload cfg3 > gpurun_out/r2_m_cfg3_head_$rep.log 2>&1
$B --workload cfg4 > gpurun_out/r2_m_cfg4_$rep.log 2>&1
SSR_B200_EARLY_TRIGGER=0 $B --workload cfg4 > gpurun_out/r2_m_cfg4_notrig_$rep.log 2>&1
SSR_B200_LIB=$OLD $B --workload cfg4 > gpurun_out/r2_m_cfg4_head_$rep.log 2>&1
done
$B --workload cfg2 > gpurun_out/r2_m_cfg2.log 2>&1
SSR_B200_LIB=$OLD $B --workload cfg2 > gpurun_out/r2_m_cfg2_head.log 2>&1
$B --outputs 64 > gpurun_out/r2_m_shard64.log 2>&1
SSR_B200_EARLY_TRIGGER=0 $B --outputs 64 > gpurun_out/r2_m_shard64_notrig.log 2>&1
SSR_B200_LIB=$OLD $B --outputs 64 > gpurun_out/r2_m_shard64_head.log 2>&1
python scripts/summarize.py gpurun_out/r2_m_*.log
cat gpurun_out/r2_m_trace_*.log | grep -v "^$" | tail -20
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_m_*.log | head
