set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_a_tests.log 2>&1; tail -3 gpurun_out/r2_a_tests.log
B="python bench.py --no-cpu-baseline --steps 300 --warmup 50"
$B > gpurun_out/r2_a_cfg5.log 2>&1
SSR_B200_MAC_WAVES=7 $B > gpurun_out/r2_a_cfg5_waves7.log 2>&1
SSR_B200_MAC_WAVES=3 $B > gpurun_out/r2_a_cfg5_waves3.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B > gpurun_out/r2_a_cfg5_nooverlap.log 2>&1
$B --outputs 64 > gpurun_out/r2_a_shard64.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --outputs 64 > gpurun_out/r2_a_shard64_nooverlap.log 2>&1
SSR_B200_PREREDUCE_MIN=8 $B --outputs 64 > gpurun_out/r2_a_shard64_prereduce8.log 2>&1
SSR_B200_MAC_WAVES=2 $B --outputs 64 > gpurun_out/r2_a_shard64_waves2.log 2>&1
$B --workload cfg4 --steps 1000 > gpurun_out/r2_a_cfg4.log 2>&1
SSR_B200_FWD_OVERLAP=0 $B --workload cfg4 --steps 1000 > gpurun_out/r2_a_cfg4_nooverlap.log 2>&1
for f in gpurun_out/r2_a_*.log; do echo $f; grep '^{' $f | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); r=d['roofline']
    print('  value %.3f e2e %.3f pipe %s ms/step %.4f mac_ms %.4f frac %.4f share %.3f fft %.4f ifft %.4f parity %s' % (d['value'], d['e2e']['value'], d['e2e_pipelined']['value'], d['ms_per_step'], r['ms_per_launch'], r['frac'], r['mac_share_of_step'], r['fft_ms_per_step'], r['ifft_ms_per_step'], d['parity'] and d['parity']['max_err']))
"; done
