"""CPU tests of bench.py's own checking code: the float64 truth the parity phase compares
the GPU output with (pinned here against the oracle's renderers, so a wrong filter index or
comparison window in bench.py cannot masquerade as an engine error), and the `config`
object both arms print."""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

entry.load_package()
import bench  # noqa: E402
import ssr_b200.workloads as wl  # noqa: E402
from oracle import conv_oracle as co  # noqa: E402


def test_parity_truth_equals_the_generic_oracle():
    B, N, M, L, nb = 64, 3, 4, 200, 10
    env = wl.envelope(L)
    o = co.GenericRendererOracle(B, M)
    for n in range(N):
        o.add_source(np.stack([wl.synth_ir(11, n * M + m, L, env, 0.1) for m in range(M)], axis=1))
    x = np.stack([wl.synth_uniform(5, n, 0, nb * B) for n in range(N)])
    y = np.concatenate([o.process(x[:, t * B:(t + 1) * B]) for t in range(nb)], axis=1)
    irs = [[wl.synth_ir(11, n * M + m, L, env, 0.1) for n in range(N)] for m in range(M)]
    truth = bench.parity_truth(x, irs)
    assert np.abs(y[:, B:] - truth[:, B:]).max() <= 1e-6       # block 0 is the fade-in


def test_parity_truth_for_brs_holds_once_the_first_staggered_switch_is_through():
    """bench.py compares Binaural / BRS from block P on: until then the renderer crossfades
    between the filter table before and after rotate_queues() (convolver.h:618-699)."""
    B, N, A, L, nb = 64, 3, 8, 300, 12
    P = (L + B - 1) // B
    env = wl.envelope(L)
    o = co.BrsRendererOracle(B)
    for n in range(N):
        o.add_source(np.stack([wl.synth_ir(7, n * 2 * A + c, L, env, 0.1) for c in range(2 * A)], axis=1))
    x = np.stack([wl.synth_uniform(3, n, 0, nb * B) for n in range(N)])
    o.reference_orientation = 0.0
    y = np.concatenate([o.process(x[:, t * B:(t + 1) * B]) for t in range(nb)], axis=1)
    # the index formula bench.py uses for azimuth 0 (src/brsrenderer.h:147-155)
    idx = int(np.fmod(np.float32(-90.0) * np.float32(A) / np.float32(360.0) + np.float32(0.5), A) + A) % A
    irs = [[wl.synth_ir(7, n * 2 * A + 2 * idx + ear, L, env, 0.1) for n in range(N)] for ear in (0, 1)]
    truth = bench.parity_truth(x, irs)
    assert np.abs(y[:, P * B:] - truth[:, P * B:]).max() <= 1e-6
    assert np.abs(y[:, B:2 * B] - truth[:, B:2 * B]).max() > 1e-3   # ... and not before


def test_both_arms_print_the_same_config_object():
    for name in ("cfg4", "cfg5", "cfg3"):
        w = copy.copy(wl.WORKLOADS[name])
        for world in (1, 2, 8):
            if w.kind != "generic" and world > 1:
                continue
            a, l2a = bench.make_config(w, False, world, w.kind != "generic")
            b, l2b = bench.make_config(w, False, world, w.kind != "generic")
            assert a == b and l2a == l2b
            assert a["outputs_per_gpu"] == (w.outputs // world if w.kind == "generic" else 2)
    cfg, l2 = bench.make_config(wl.WORKLOADS["cfg4"], False, 8, False)
    assert l2 and "L2-RESIDENT" in cfg["l2"]           # 33.5 MB of spectra per GPU
    cfg, l2 = bench.make_config(wl.WORKLOADS["cfg5"], False, 8, False)
    assert not l2 and cfg["l2"].startswith("inputs>L2")


def test_reference_arm_never_needs_the_product_library():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ref_part = src[src.index("# ----------------------------------------------------------------- reference arm"):
                   src.index("# ----------------------------------------------------------------- parity")]
    ref_part = ref_part.replace('"ssr_b200.capi" in sys.modules', "")   # (the line that REPORTS whether it was loaded)
    assert "capi" not in ref_part, "the reference arm must generate its data with the numpy twins (workloads.py)"


def test_mac_traffic_table_is_reproducible_from_the_committed_ncu_pages():
    """profiles/mac_traffic.json (what bench.py reports as roofline.traffic) is what profiles/traffic_from_ncu.py
    extracts from the committed raw pages of the ncu captures: same DRAM bytes per MAC launch, same grid."""
    import importlib.util
    import json
    prof = os.path.join(ROOT, "profiles")
    spec = importlib.util.spec_from_file_location("traffic_from_ncu", os.path.join(prof, "traffic_from_ncu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    table = json.load(open(os.path.join(prof, "mac_traffic.json")))
    pages = {"cfg5_gpus1": "r2c_ncu/r2b_mac_cfg5_gpus1.raw.csv", "cfg5_gpus8": "r2c_ncu/r2b_mac_cfg5_gpus8.raw.csv",
             "cfg4_gpus1": "r2c_ncu/r2b_mac_cfg4_gpus1.raw.csv", "cfg3_gpus1": "r2b_ncu/r2b_cfg3.raw.csv",
             "cfg2_gpus1": "r2b_ncu/r2b_cfg2.raw.csv"}
    for key, rel in pages.items():
        e = mod.entry(os.path.join(prof, rel))
        assert e is not None, key
        assert int(e["grid"]) == int(table[key]["grid"]), key
        assert abs(e["bytes"] - table[key]["bytes"]) <= 1e-6 * table[key]["bytes"], key
        assert e["kernel"] == table[key]["kernel"], key
    # the judged kernel's traffic stays within 1 % of the algorithmic bytes on one GPU
    algorithmic = 8192 * (47 * 128 * 512 + 128 * 47 + 512)
    assert 1.0 <= table["cfg5_gpus1"]["bytes"] / algorithmic <= 1.01
