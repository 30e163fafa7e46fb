"""profiles/conv_tc_traffic.json (read by bench.py for roofline.traffic) from an ncu dram-bytes launch list of the conv kernels:
   python scripts/conv_traffic.py gpurun_out/conv_dram_<tag>.csv profiles/conv_dram_<tag>.csv"""
import csv, json, sys, collections
src, kept = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    if "dram__bytes" not in r["Metric Name"]: continue
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"].lower()
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    per[r["ID"]] = per.get(r["ID"], 0.0) + v
ids = list(per)
n = 61
last = ids[-n:]                      # the conv launches of the last U-Net step
tot = sum(per[i] for i in last)
open(kept, "w").writelines(lines)
json.dump({"avg_dram_bytes_per_launch": tot / n, "launches": n, "dram_mb_per_unet_step": tot / 1e6,
           "source": f"{kept}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc over scripts/step_n.py 64 2, "
                     f"mean over the {n} conv_tc launches of the last U-Net step (B=64)"},
          open("profiles/conv_tc_traffic.json", "w"), indent=1)
print("launch ids", len(ids), "-> per step MB", tot / 1e6, "avg per launch MB", tot / n / 1e6)
