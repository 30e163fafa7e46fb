import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
byop = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) <= iN: continue
    try: n = int(r[iN])
    except ValueError: continue
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0]
    e = byop[op]; e[0] += n; e[1] += int(r[iE] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: e[2][hdr[i]] += v
tot = sum(e[0] for e in byop.values())
print("total samples", tot)
for op, e in sorted(byop.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{op:12s} {100*e[0]/tot:5.1f}%  exec={e[1]:8d}  {e[2].most_common(3)}")
