#!/bin/bash
# DRAM bytes of every launch of the sampling step with the caches left alone between kernels (ncu --cache-control none):
# how much of the activation traffic actually reaches HBM.   usage: bash scripts/gpu_dram_step.sh <tag> [batch]
cd "$(dirname "$0")/.."
tag=${1:-r02}; b=${2:-64}
mkdir -p gpurun_out
timeout 300 python scripts/step_n.py $b 3 > gpurun_out/plain_dram_$tag.log 2>&1 || { echo plain run failed; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --csv --log-file gpurun_out/dram_step_$tag.csv python scripts/step_n.py $b 3 > gpurun_out/ncu_dram_step_$tag.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv, collections, re
lines=[l for l in open("gpurun_out/dram_step_$tag.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
per=collections.OrderedDict()
for r in rows:
    d=per.setdefault(r["ID"], {"name": re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","").replace("dd::","")})
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"].lower()
    if "dram" in r["Metric Name"]:
        v*= {"byte":1,"kbyte":1e3,"mbyte":1e6,"gbyte":1e9}.get(u,1); d["rd" if "read" in r["Metric Name"] else "wr"]=v
    else:
        d["us"]= v/1000 if u.startswith("n") else v
ids=list(per)
ims=[i for i in ids if "im2col" in per[i]["name"] or "nchw_to_nhwc_pad" in per[i]["name"]]
step=ids[ids.index(ims[-2]):ids.index(ims[-1])]
rd=sum(per[i].get("rd",0) for i in step); wr=sum(per[i].get("wr",0) for i in step); us=sum(per[i].get("us",0) for i in step)
print(f"one step: {len(step)} launches, DRAM read {rd/1e6:.1f} MB, write {wr/1e6:.1f} MB, kernel time (serialised) {us:.1f} us -> {(rd+wr)/1e6/ (us/1e6) /1e6:.2f} TB/s")
agg=collections.defaultdict(lambda:[0,0.0,0.0,0.0])
for i in step:
    a=agg[per[i]["name"][:44]]; a[0]+=1; a[1]+=per[i].get("rd",0); a[2]+=per[i].get("wr",0); a[3]+=per[i].get("us",0)
for k,a in sorted(agg.items(), key=lambda kv:-(kv[1][1]+kv[1][2])): print(f"{k:44s} n={a[0]:3d} rd {a[1]/1e6:8.1f} MB wr {a[2]/1e6:8.1f} MB {a[3]:8.1f} us")
PY
