"""Prints the key numbers of bench.py JSON lines: python profiles/summarize.py file..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:  # noqa: BLE001
        print(path, "unreadable:", exc)
        continue
    r = d.get("roofline", {})
    print(f"{path}: gpus={d.get('n_gpus')} value={d.get('value'):.3f} e2e={d.get('e2e', {}).get('value', 0):.3f} "
          f"ms/step={d.get('ms_per_step'):.4f} p99={d.get('p99_block_ms', 0):.3f} "
          f"mac={r.get('achieved', 0):.0f}GB/s frac={r.get('frac', 0):.3f} mac_ms={r.get('ms_per_launch', 0):.4f} "
          f"fft_ms={r.get('fft_ms_per_step', 0):.4f} ifft_ms={r.get('ifft_ms_per_step', 0):.4f} "
          f"clk={d.get('clocks')} cpu={d.get('cpu_baseline', {}).get('value')}")
