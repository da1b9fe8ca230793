"""Rebuilds profiles/mac_traffic.json entries from the raw CSV pages of the ncu captures.

    python profiles/traffic_from_ncu.py <dir with r2b_*.raw.csv> <commit> [--write]

Every `r2b_<tag>.raw.csv` is `ncu -i r2b_<tag>.ncu-rep --page raw --csv` of a `--set full --clock-control none`
capture made by profiles/ncu_capture_r2b.sh.  The MAC launches of a capture (kernel name contains `mac_`) are
averaged: DRAM bytes read + written per launch, duration under ncu, DRAM throughput, grid.  bench.py reports the
`bytes` of the entry `<workload>_gpus<N>` as roofline.traffic only when its own plan launched the same grid.
"""
import csv
import json
import os
import sys

TAGS = {  # capture tag -> key of mac_traffic.json
    "mac_cfg5_gpus1": "cfg5_gpus1", "mac_cfg5_gpus2": "cfg5_gpus2", "mac_cfg5_gpus4": "cfg5_gpus4",
    "mac_cfg5_gpus8": "cfg5_gpus8", "mac_cfg4_gpus1": "cfg4_gpus1", "mac_cfg4_gpus8": "cfg4_gpus8",
    "cfg3": "cfg3_gpus1", "cfg2": "cfg2_gpus1",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3,
        "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}


def num(s):
    return float(s.replace(",", "")) if s not in ("", "n/a") else 0.0


def entry(path):
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                     "dram__bytes_write.sum", "launch__grid_size",
                                     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    mac = [r for r in body if "mac_" in r[col["Kernel Name"]]]
    if not mac:
        return None
    def scaled(r, k):
        return num(r[col[k]]) * UNIT.get(units[col[k]], 1.0)
    n = len(mac)
    return {
        "bytes": sum(scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum") for r in mac) / n,
        "launches": n,
        "grid": int(num(mac[0][col["launch__grid_size"]])),
        "kernel": mac[0][col["Kernel Name"]],
        "ms_under_ncu": sum(scaled(r, "gpu__time_duration.sum") for r in mac) / n,
        "dram_throughput_pct": sum(num(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) for r in mac) / n,
    }


def main():
    src, commit = sys.argv[1], sys.argv[2]
    write = "--write" in sys.argv
    here = os.path.dirname(os.path.abspath(__file__))
    tpath = os.path.join(here, "mac_traffic.json")
    table = json.load(open(tpath))
    for tag, key in TAGS.items():
        p = os.path.join(src, f"r2b_{tag}.raw.csv")
        if not os.path.exists(p):
            continue
        e = entry(p)
        if e is None:
            print(f"{tag}: no MAC launch in the capture")
            continue
        e["commit"] = commit
        e["capture"] = f"profiles/ncu_capture_r2b.sh -> r2b_{tag}; ncu --set full --clock-control none"
        old = table.get(key, {})
        print(f"{key}: {e['kernel'][:50]} grid {e['grid']} bytes {e['bytes']:.0f} ({e['ms_under_ncu']:.4f} ms, "
              f"{e['dram_throughput_pct']:.1f} % DRAM)   was: grid {old.get('grid')} bytes {old.get('bytes')}")
        table[key] = e
    if write:
        note = table.pop("_note", None)
        if note:
            table["_note"] = note
        json.dump(table, open(tpath, "w"), indent=1)
        print("written", tpath)


if __name__ == "__main__":
    main()
