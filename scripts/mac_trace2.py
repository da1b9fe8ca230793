"""Per-CTA timeline of the bulk-copy MACs (development): python scripts/mac_trace2.py cfg4|cfg3|cfg5 [outputs]
Run with SSR_B200_TRACE_ISSUE=1: per CTA {start, first stage complete, end, first bulk copy issued}."""
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

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = wl.WORKLOADS[name]
outputs = int(sys.argv[2]) if len(sys.argv) > 2 else w.outputs
env = wl.envelope(w.taps); scale = wl.ir_scale(w.sources, env)
eng = capi.Engine(w.block, w.fs, device=0)
if w.kind == "generic":
    ren = capi.Renderer(eng, capi.GENERIC, outputs=outputs)
    for n in range(w.sources):
        ren.add_source_synthetic(w.taps, outputs, w.ir_seed, channel_id_offset=n * outputs, envelope=env, scale=scale)
else:
    ren = capi.Renderer(eng, capi.BRS, outputs=2)
    for n in range(w.sources):
        ren.add_source_synthetic(w.taps, w.channels, w.ir_seed, channel_id_offset=n * w.channels, envelope=env, scale=scale)
    outputs = 2
x = torch.from_numpy(np.stack([wl.signal_block(w, t) for t in range(4)])).cuda()
y = torch.zeros((outputs, w.block), device="cuda")
torch.cuda.synchronize()
az = 0.0
for rep in range(6):
    for t in range(60):
        if w.kind != "generic":
            az += 1.0
            ren.set_reference_orientation(az)
        ren.process_device(x[t % 4].data_ptr(), y.data_ptr())
    eng.synchronize()
    G = eng.stats()["last_mac_grid"]
    tr = eng.mac_trace(G).astype(np.int64)
    t0 = tr[:, 0].min()
    start, first, end, issue = [(tr[:, i] - t0) / 1e3 for i in (0, 1, 2, 3)]
    q = lambda a: f"{a.min():.1f} / {np.median(a):.1f} / {a.max():.1f}"
    print(f"{name} grid {G}: start {q(start)} | first issue {q(issue)} | first stage {q(first)} | end {q(end)} us (min / median / max)")
    if w.kind != "generic":
        full = eng.mac_trace(4096).reshape(-1)
        dbg = full[8192:8192 + 8 * 32].reshape(32, 8).astype(np.int64)
        live = dbg[:, 0] > 0
        if live.any():
            d = dbg[live]
            base = d[:, 0].min()
            names = ["start", "dep wait", "sync1", "reduce", "sync2", "fft done", "stored"]
            for i, nm in enumerate(names):
                col = d[:, i][d[:, i] > 0]
                if len(col):
                    print(f"   inverse cluster {nm:9s}: {(col.min() - base) / 1e3:6.1f} .. {(col.max() - base) / 1e3:6.1f} us   (MAC start -> {(col.min() - t0) / 1e3:6.1f})")
