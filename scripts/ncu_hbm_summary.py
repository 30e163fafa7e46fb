"""Per-kernel summary of an ncu --csv metrics pass (time, DRAM bytes, L2 bytes): python scripts/ncu_hbm_summary.py a.csv [b.csv ...]"""
import collections, csv, re, sys


def val(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
             "Gbyte": 1e9, "%": 1.0}
    return v * scale.get(u, 1.0)


for path in sys.argv[1:]:
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
    per = collections.OrderedDict()
    for r in rows:
        key = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dd::", ""), r["Grid Size"])
        per.setdefault(key, {})[r["Metric Name"]] = val(r)
    agg = collections.OrderedDict()
    for (_, name, grid), m in per.items():
        a = agg.setdefault((name, grid), dict(n=0, t=0.0, rd=0.0, wr=0.0, l2=0.0, pct=0.0))
        a["n"] += 1; a["t"] += m.get("gpu__time_duration.sum", 0.0); a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0); a["l2"] += m.get("lts__t_bytes.sum", 0.0)
        a["pct"] += m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    print(path)
    print(f"{'kernel':44s} {'grid':>16s} {'n':>3s} {'us':>8s} {'DRAM MB':>9s} {'L2 MB':>9s} {'DRAM GB/s':>10s} {'L2 GB/s':>9s} {'dram %pk':>8s}")
    for (name, grid), a in agg.items():
        n, t = a["n"], a["t"] / a["n"]
        dram, l2 = (a["rd"] + a["wr"]) / n, a["l2"] / n
        print(f"{name[:44]:44s} {grid:>16s} {n:3d} {t:8.2f} {dram / 1e6:9.2f} {l2 / 1e6:9.2f} {dram / t / 1e3:10.0f} {l2 / t / 1e3:9.0f} {a['pct'] / n:8.1f}")
