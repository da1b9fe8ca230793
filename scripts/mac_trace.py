"""Per-CTA timeline of the bulk-copy MAC (development): python scripts/mac_trace.py [outputs]
Prints, every 12 blocks, when the CTAs of the last MAC launch started and ended -- before and after the
share calibration (Renderer::calibrate) re-cuts the shares at blocks 24 and 48."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
import ssr_b200.capi as capi  # noqa: E402
import ssr_b200.workloads as wl  # noqa: E402
import torch  # noqa: E402

outputs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = wl.WORKLOADS["cfg5"]
env = wl.envelope(w.taps); scale = wl.ir_scale(w.sources, env)
eng = capi.Engine(w.block, w.fs, device=0)
ren = capi.Renderer(eng, capi.GENERIC, outputs=outputs)
for n in range(w.sources):
    ren.add_source_synthetic(w.taps, outputs, w.ir_seed, channel_id_offset=n * outputs, envelope=env, scale=scale)
x = torch.from_numpy(np.stack([wl.signal_block(w, t) for t in range(4)])).cuda()
y = torch.zeros((outputs, w.block), device="cuda")
torch.cuda.synchronize()
for rep in range(8):
    for t in range(12):
        ren.process_device(x[t % 4].data_ptr(), y.data_ptr())
    eng.synchronize()
    tr = eng.mac_trace(148).astype(np.int64)
    t0 = tr[:, 0].min()
    start, first, end = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3
    print(f"block {12 * (rep + 1)}: CTA start spread {start.max():.1f} us | first data at {first.min():.1f} .. {first.max():.1f} us | "
          f"end {end.min():.1f} .. {end.max():.1f} us (median {np.median(end):.1f}, p10 {np.percentile(end, 10):.1f}, "
          f"p90 {np.percentile(end, 90):.1f}) | per-CTA busy {np.median(end - start):.1f} us")
    order = np.argsort(end)
    print("   slowest CTAs:", [(int(i), round(float(end[i]), 1)) for i in order[-6:]], " fastest:", [(int(i), round(float(end[i]), 1)) for i in order[:4]])
