"""Top warp-stall lines per kernel from `ncu -i rep --page source --csv`:
    python profiles/ncu_source_top.py source.csv [kernel-substring] [lines]"""
import csv
import sys

csv.field_size_limit(10**9)
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 12
kernels, cur = [], None
for r in csv.reader(open(path)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
seen = set()
for k in kernels:
    if want not in k["name"] or k["hdr"] is None:
        continue
    h = k["hdr"]
    si, ai = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ai] or 0) for r in k["rows"])
    key = (k["name"], tot, len(k["rows"]))
    if key in seen:
        continue
    seen.add(key)
    stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    print(f"### `{k['name'][:70]}` — {tot} stall samples, {len(k['rows'])} SASS lines\n")
    by_reason = sorted(((sum(int(r[i] or 0) for r in k["rows"]), h[i]) for i in stall), reverse=True)[:6]
    print("by reason: " + ", ".join(f"{n} {c} ({100 * n / max(1, tot):.0f} %)" for n, c in by_reason if n) + "\n")
    print("| samples | share | SASS | main reason |\n|---|---|---|---|")
    for r in sorted(k["rows"], key=lambda r: -int(r[ai] or 0))[:nlines]:
        n = int(r[ai] or 0)
        why = max(((int(r[i] or 0), h[i]) for i in stall))[1]
        print(f"| {n} | {100 * n / max(1, tot):.1f} % | `{r[si].strip()[:80]}` | {why} |")
    print()
