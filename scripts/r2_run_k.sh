set -x
nproc; free -g | head -2
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2_k_reference_cfg5.log 2>&1; tail -4 gpurun_out/r2_k_reference_cfg5.log | cut -c1-900
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_k_bench_driver_style.log 2>&1
python bench.py > gpurun_out/r2_k_bench_default.log 2>&1
for w in cfg2 cfg3 cfg4; do python bench.py --workload $w > gpurun_out/r2_k_bench_$w.log 2>&1; done
ssr-pre-0.4.0_b200/lib/perf_static_convolver 2000 > gpurun_out/r2_k_cfg1.log 2>&1; tail -3 gpurun_out/r2_k_cfg1.log
python scripts/summarize.py gpurun_out/r2_k_bench_*.log
python - <<'PY'
import json
for f in ["driver_style","default","cfg2","cfg3","cfg4"]:
    for ln in open(f"gpurun_out/r2_k_bench_{f}.log"):
        if ln.startswith("{"):
            d=json.loads(ln); print(f, "cpu_baseline", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"), str(d.get("cpu_baseline",{}).get("sample"))[:160])
PY
python -c "import __graft_entry__ as g; g.smoke()"
