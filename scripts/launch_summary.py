"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares,
optionally per NVTX range. Usage: python scripts/launch_summary.py launches.csv "<header line>" """
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
nv = [i for i, h in enumerate(hdr) if "Push/Pop" in h]
nv = nv[0] if nv else None


def us(r):
    v = float(r[mv].replace(",", ""))
    return v / 1e3 if r[mu] in ("ns", "nsecond") else v * 1e3 if r[mu] in ("ms", "msecond") else v


def short(name):
    name = name.replace("void ", "").replace("mmrec::<unnamed>::", "").replace("at::native::", "at::")
    return name[:name.index("(")] if "(" in name else name


if len(sys.argv) > 2:
    print("# " + sys.argv[2])
groups = collections.OrderedDict()
for r in data:
    rng = "all"
    if nv is not None:
        rng = "timed_eval" if "timed_eval" in r[nv] else "timed_steps" if "timed_steps" in r[nv] else "other"
    groups.setdefault(rng, []).append(r)
for rng, rs in groups.items():
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rs:
        a = agg[short(r[kn])[:100]]
        a[0] += 1
        a[1] += us(r)
    tot = sum(v[1] for v in agg.values())
    print(f"## range {rng}: {len(rs)} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare SHARES)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:4d}  avg {t / c:8.1f} us  {k}")
