"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths
(tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mbarrier -> SYNCS), from
`cuobjdump -sass recombiner_b200/librecombiner_b200.so`.  Usage: python profiles/scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "recombiner_b200", "librecombiner_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.SIN", "MUFU.COS", "DFMA",
             "HMMA", "FFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for mn in MNEMONICS:
        if op == mn or op.startswith(mn + "."):
            counts[cur][mn] += 1
    counts[cur]["_total"] += 1
print("# cuobjdump -sass of librecombiner_b200.so (sm_100a), instruction counts per kernel")
print("%-78s %7s " % ("kernel", "instr") + " ".join("%8s" % m for m in MNEMONICS))
tot = collections.Counter()
for k, c in counts.items():
    print("%-78s %7d " % (k[:78], c["_total"]) + " ".join("%8d" % c[m] for m in MNEMONICS))
    tot.update(c)
print("%-78s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[m] for m in MNEMONICS))
