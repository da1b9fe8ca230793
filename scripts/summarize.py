"""Prints one line per bench.py JSON line found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    for ln in open(f):
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        r = d.get("roofline") or {}
        par = d.get("parity")
        pipe = (d.get("e2e_pipelined") or {}).get("value")
        print("%-52s value %8.3f e2e %8.3f pipe %s ms/step %.4f mac_ms %.4f frac %.4f share %.3f fft %.4f ifft %.4f parity %s clk %s %s" % (
            f.split("/")[-1], d["value"], d["e2e"]["value"], "%8.3f" % pipe if pipe else "   -    ", d["ms_per_step"],
            r.get("ms_per_launch", 0), r.get("frac", 0), r.get("mac_share_of_step", 0), r.get("fft_ms_per_step", 0),
            r.get("ifft_ms_per_step", 0), ("%.2e" % par["max_err"]) if par else None,
            (d.get("clocks") or {}).get("sm_mhz"), ",".join((d.get("clocks") or {}).get("reasons") or [])))
