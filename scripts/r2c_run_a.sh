#!/bin/bash
# block overlap A/B (one B200): tests of the renderers / gather / full size, then bench lines with and without
OUT=gpurun_out/r2c_a; mkdir -p $OUT
python -m pytest tests/test_gpu_renderers.py tests/test_gpu_gather.py tests/test_gpu_hostgather.py tests/test_gpu_fullsize.py tests/test_gpu_nonuniform.py -q -x 2>&1 | tail -6
B="python bench.py --no-cpu-baseline --no-sustained"
for ov in 1 0; do
  SSR_B200_BLOCK_OVERLAP=$ov $B --outputs 64 --steps 1000 --warmup 100 > $OUT/shard64_ov$ov.json 2> $OUT/shard64_ov$ov.err; echo "shard64 ov=$ov rc=$?"
  SSR_B200_BLOCK_OVERLAP=$ov $B --workload cfg4 --steps 2000 --warmup 200 > $OUT/cfg4_ov$ov.json 2> $OUT/cfg4_ov$ov.err; echo "cfg4 ov=$ov rc=$?"
  SSR_B200_BLOCK_OVERLAP=$ov $B --workload cfg4 --outputs 8 --steps 2000 --warmup 200 > $OUT/cfg4s8_ov$ov.json 2> $OUT/cfg4s8_ov$ov.err; echo "cfg4 shard8 ov=$ov rc=$?"
done
python profiles/summarize.py $OUT/*.json 2>/dev/null | cut -c1-230
