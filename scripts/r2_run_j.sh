export SSR_B200_MAC_CALIBRATE=0
for o in 0 1 2; do
  echo "== term order $o, shard64"; SSR_B200_TERM_ORDER=$o python scripts/mac_trace.py 64 2>&1 | grep block | tail -3
done
for o in 0 1 2; do
  echo "== term order $o, full cfg5"; SSR_B200_TERM_ORDER=$o python scripts/mac_trace.py 512 2>&1 | grep block | tail -2
done
B="python bench.py --no-cpu-baseline --steps 1000 --warmup 100 --no-parity"
for o in 0 1 2; do SSR_B200_TERM_ORDER=$o $B --outputs 64 > gpurun_out/r2_j_shard64_order$o.log 2>&1; done
python scripts/summarize.py gpurun_out/r2_j_*.log
