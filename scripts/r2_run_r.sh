set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_r_tests.log 2>&1; tail -5 gpurun_out/r2_r_tests.log
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100"
$B --nonuniform 1x4,2x2,4x2,8x4 > gpurun_out/r2_r_nu4_cfg5.log 2>&1
$B --nonuniform 1x4,2x2,4x2,8x4 --outputs 64 > gpurun_out/r2_r_nu4_shard64.log 2>&1
$B --nonuniform 1x4,2x2,4x10 --outputs 64 > gpurun_out/r2_r_nu3_shard64.log 2>&1
$B --nonuniform 1x8,8x5 > gpurun_out/r2_r_nu2b_cfg5.log 2>&1
for f in gpurun_out/r2_r_nu*.log; do grep -h "^{" $f | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('$f', round(d['value'],2), round(d['e2e']['value'],2), d['ms_per_step'], d['parity']['max_err'], d['roofline']['frac'], d['roofline']['bytes_saved_factor'], d['clocks']['sm_mhz'])
"; done
grep -h "Error\|error\|Traceback\|PARITY" gpurun_out/r2_r_*.log | head
