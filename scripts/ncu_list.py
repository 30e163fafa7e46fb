"""Summarise an ncu --csv gpu__time_duration launch list: per-kernel totals and one step in launch order."""
import csv, re, sys, collections
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
def nm(r): return re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dd::", "")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    agg[nm(r)][0] += 1; agg[nm(r)][1] += us(r)
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)} total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:48]:48s} n={v[0]:4d} sum={v[1]:9.1f} avg={v[1] / v[0]:7.2f} share={100 * v[1] / tot:5.1f}%")
if len(sys.argv) > 2:
    idx = [i for i, r in enumerate(rows) if "im2col" in r["Kernel Name"] or "nchw_to_nhwc_pad" in r["Kernel Name"]]
    if len(idx) >= 2:
        for r in rows[idx[-2]:idx[-1]]:
            print(f"{nm(r)[:30]:30s} grid={r['Grid Size']:>16s} {us(r):8.2f} us")
