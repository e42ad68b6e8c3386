"""Instruction histogram of every kernel in libvsf_cuda.so (cuobjdump -sass): the mnemonics that
prove what the code is made of (UTCIMMA / UTCQMMA = tcgen05.mma kind::i8 / kind::f8f6f4, LDTM =
tcgen05.ld, UBLKCP = TMA 1-D bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier, POPC / LOP3 =
the integer-pipe engine, ...).  No GPU needed.  Writes profiles/<tag>_sass_histogram.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "vision_slam_frontend_b200", "libvsf_cuda.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur[op.split(".")[0]] += 1
        cur["_total"] += 1
KEY = ["UTCIMMA", "UTCQMMA", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "POPC", "LOP3", "VIMNMX",
       "VIMNMX3", "SHFL", "VOTE", "MATCH", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "BAR", "DFMA", "DMUL", "DADD",
       "MUFU", "IMAD", "IADD3", "LDL", "STL", "ACQBULK", "ELECT", "ERRBAR", "NANOSLEEP"]
path = os.path.join(ROOT, "profiles", tag + "_sass_histogram.txt")
with open(path, "w") as f:
    f.write("# cuobjdump -sass libvsf_cuda.so (sm_100a), instruction counts per kernel; only mnemonics that occur are listed\n")
    tot = collections.Counter()
    for name, c in kernels.items():
        tot.update(c)
        f.write("\n%s  (%d instructions)\n" % (name, c["_total"]))
        f.write("   " + "  ".join("%s %d" % (k, c[k]) for k in KEY if c[k]) + "\n")
    f.write("\nWHOLE LIBRARY\n   " + "  ".join("%s %d" % (k, tot[k]) for k in KEY if tot[k]) + "\n")
print(open(path).read()[-1500:])
