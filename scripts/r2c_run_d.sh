#!/bin/bash
# Binaural / BRS block overlap (one B200): dynamic-renderer tests, then cfg3 / cfg2 bench lines with and without
OUT=gpurun_out/r2c_d; mkdir -p $OUT
python -m pytest tests/test_gpu_renderers.py tests/test_gpu_golden.py tests/test_gpu_control_thread.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_cpp.py -q -x 2>&1 | tail -8
B="python bench.py --no-cpu-baseline --no-sustained"
for ov in 1 0; do
  SSR_B200_DYN_BLOCK_OVERLAP=$ov $B --workload cfg3 --steps 3000 --warmup 300 > $OUT/cfg3_ov$ov.json 2> $OUT/cfg3_ov$ov.err; echo "cfg3 ov=$ov rc=$?"
  SSR_B200_DYN_BLOCK_OVERLAP=$ov $B --workload cfg2 --steps 3000 --warmup 300 > $OUT/cfg2_ov$ov.json 2> $OUT/cfg2_ov$ov.err; echo "cfg2 ov=$ov rc=$?"
done
python profiles/summarize.py $OUT/*.json 2>/dev/null | cut -c1-230
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_d/*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], 'pipe', (d.get('e2e_pipelined') or {}).get('value'), 'p99', d.get('p99_block_ms'), 'parity', d['parity']['max_err'], 'devclk', (d['roofline'].get('device_clock') or {}).get('ms'))
PY
tail -3 $OUT/*.err | head -30
