"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page):  python scripts/ncu_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iSrc = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) <= iN: continue
    try: n = int(r[iN])
    except ValueError: continue
    data.append((n, k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, k, r in sorted(data, reverse=True)[:N]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100*n/max(tot,1):5.1f}% {r[iS].strip()[:70]:70s} exec={r[iSrc]:>7s} {st}")
agg = {}
for n, k, r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True))
