set -x
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
for rep in 1 2; do
$B --workload cfg4 > gpurun_out/r2_q_cfg4_$rep.log 2>&1
SSR_B200_INV_TRIGGER=1 $B --workload cfg4 > gpurun_out/r2_q_cfg4_invtrig_$rep.log 2>&1
$B --outputs 64 > gpurun_out/r2_q_shard64_$rep.log 2>&1
SSR_B200_INV_TRIGGER=1 $B --outputs 64 > gpurun_out/r2_q_shard64_invtrig_$rep.log 2>&1
done
$B --workload cfg3 > gpurun_out/r2_q_cfg3.log 2>&1
$B --nonuniform 1x4,2x2,4x10 > gpurun_out/r2_q_nu3_cfg5.log 2>&1
SSR_B200_INV_TRIGGER=1 $B --nonuniform 1x4,2x2,4x10 > gpurun_out/r2_q_nu3_cfg5_invtrig.log 2>&1
python scripts/summarize.py gpurun_out/r2_q_*.log
grep -h "device_clock" gpurun_out/r2_q_cfg3.log gpurun_out/r2_q_cfg4_1.log gpurun_out/r2_q_shard64_1.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['config']['workload'][:40], d['roofline']['device_clock'], d['roofline']['ms_per_launch'])
"
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_q_*.log | head
