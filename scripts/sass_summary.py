"""Per-kernel SASS opcode summary of libddb200.so (run where cuobjdump is: no GPU needed).
   python scripts/sass_summary.py > profiles/sass_opcodes_<tag>.txt
Counts the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, REDG / ATOMG = global atomics,
STG.*256 / LDG.*256 = 32-byte global accesses, MUFU = special-function unit."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "downsampled_diffusion_b200", "libddb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), {}
KEYS = [("UTC.MMA", r"^UTC[A-Z]*MMA"), ("UTCBAR", r"^UTCBAR"), ("LDTM", r"^LDTM"), ("UTMALDG", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"), ("SYNCS", r"^SYNCS"),
        ("HMMA", r"^HMMA"), ("MUFU", r"^MUFU"), ("REDG/ATOMG", r"^(REDG|ATOMG|RED)"), ("LDG.256", r"^LDG.*\.256"), ("STG.256", r"^STG.*\.256"), ("BAR", r"^BAR"),
        ("UCGABAR", r"^UCGABAR"), ("LDS", r"^LDS"), ("STS", r"^STS")]
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("dd::", "")
        counts[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for name, rx in KEYS:
            if re.match(rx, op):
                counts[kern][name] += 1
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{k[0]:>10s}" for k in KEYS))
for k, c in counts.items():
    if total[k] == 0:
        continue
    print(f"{k[:64]:64s} {total[k]:6d} " + " ".join(f"{c.get(n[0], 0):10d}" for n in KEYS))
