#!/usr/bin/env python
"""SASS evidence of the in-tree library: counts of the Blackwell mnemonics per kernel.

    python tools/sass_summary.py [maskunet_b200/libmaskunet_b200.so] > profiles/r02_sass_summary.txt
"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "maskunet_b200/libmaskunet_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2"]
per, cur, idx = collections.OrderedDict(), None, 0
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        full = names[idx].replace("(int)", "").replace("(bool)", ""); idx += 1
        cur = full[:full.index(">(") + 1] if ">(" in full else re.sub(r"\(.*", "", full)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for k in KEYS:
        if op == k or op.startswith(k + "."):
            per[cur][k] += 1
    if op.startswith("UTCHMMA") and "2CTA" in line and not (op == "UTCHMMA.2CTA" or op.startswith("UTCHMMA.2CTA.")):
        per[cur]["UTCHMMA.2CTA"] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS evidence, libmaskunet_b200.so (cuobjdump -sass, sm_100a), round 2, final build")
print("# tcgen05.mma -> UTCHMMA (cta_group::2: .2CTA), tcgen05.commit -> UTCBAR, tcgen05.ld/st -> LDTM/STTM, TMA load/store/reduce ->")
print("# UTMALDG/UTMASTG/UTMAREDG, cp.async.bulk -> UBLKCP; legacy mma.sync (HMMA) count must be 0")
print("totals: " + " ".join(f"{k}={tot[k]}" for k in KEYS))
print()
print(f"{'kernel':100s} " + " ".join(f"{k:>8s}" for k in KEYS[:9]))
for name, c in sorted(per.items(), key=lambda kv: -kv[1]["UTCHMMA"]):
    if c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"]:
        print(f"{name[:100]:100s} " + " ".join(f"{c[k]:8d}" for k in KEYS[:9]))
