"""Throughput of filter-set loading (forward FFT at load time): transforms per second and
GFLOP/s by the 56 320-flop convention.  python profiles/fft_load_rate.py  (needs a GPU)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
import ssr_b200.capi as capi  # noqa: E402
import ssr_b200.workloads as wl  # noqa: E402

B, taps, channels = 1024, 48000, 8192
eng = capi.Engine(B)
env = wl.envelope(taps)
fs = capi.FilterSet(eng, channels, capi.min_partitions(B, taps))
fs.generate(taps, seed=1, envelope=env, scale=0.01)  # warm-up
eng.synchronize()
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    fs.generate(taps, seed=1, envelope=env, scale=0.01)
    eng.synchronize()
    best = min(best, time.perf_counter() - t0)
n = channels * fs.partitions
print(f"fft mode={os.environ.get('SSR_B200_FFT', 'fft1024')}: {n} transforms in {best * 1e3:.3f} ms = "
      f"{n / best / 1e6:.1f} M transforms/s = {n * 56320 / best / 1e12:.2f} TFLOP/s, "
      f"{n * 8192 / best / 1e9:.0f} GB/s written")
