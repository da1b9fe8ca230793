set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_g_tests.log 2>&1; tail -5 gpurun_out/r2_g_tests.log
B="python bench.py --no-cpu-baseline --steps 300 --warmup 50"
$B > gpurun_out/r2_g_cfg5.log 2>&1
$B --workload cfg4 --steps 1000 > gpurun_out/r2_g_cfg4.log 2>&1
python scripts/summarize.py gpurun_out/r2_g_c*.log
bash profiles/ncu_capture_r2.sh 2>&1 | tail -25
