"""Per-kernel SASS opcode counts of libmmrec_b200.so (cuobjdump -sass): which kernels use the
Blackwell-native instructions (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG =
cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier) and which run on mma.sync (HMMA).
    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "recommendar-systems_b200", "libmmrec_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU", "RED", "ATOM",
         "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        counts[cur][op] += 1
names = list(counts)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("# kernel | total instructions | " + " ".join(WATCH))
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    c = counts[n]
    short = re.sub(r"\(anonymous namespace\)::", "", d)
    short = re.sub(r"\(.*", "", short)[:90]
    print(f"{short:92s} {sum(c.values()):7d} | " + " ".join(f"{k}={c[k]}" for k in WATCH if c[k]))
