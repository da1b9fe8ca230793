#!/bin/bash
# control-upload kernel without the system-scope fence / with the early counter fetch: cfg3 / cfg2 lines + dynamic tests
OUT=gpurun_out/r2c_f; mkdir -p $OUT
python -m pytest tests/test_gpu_renderers.py tests/test_gpu_golden.py tests/test_gpu_control_thread.py -q -x 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --no-sustained"
$B --workload cfg3 --steps 3000 --warmup 300 > $OUT/cfg3.json 2> $OUT/cfg3.err; echo "cfg3 rc=$?"
$B --workload cfg2 --steps 3000 --warmup 300 > $OUT/cfg2.json 2> $OUT/cfg2.err; echo "cfg2 rc=$?"
python profiles/summarize.py $OUT/*.json 2>/dev/null | cut -c1-200
