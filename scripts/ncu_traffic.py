"""profiles/rNN_ncu_traffic.json from an `ncu --page raw --csv` export of scripts/ncu_kernels.py: DRAM bytes
(read + write) of the warm (last) launch of every hot kernel, keyed the way bench.py's roofline blocks name them.
Usage: python scripts/ncu_traffic.py raw.csv out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], [r for r in rows[hi + 2:] if len(r) == len(rows[hi])]
kn = hdr.index("Kernel Name")


def val(r, m):
    i = hdr.index(m)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
                "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1}.get(u, 1)


def pick(sub, which="last", grid=None):
    c = [r for r in data if sub in r[kn] and (grid is None or int(float(r[hdr.index("launch__grid_size")])) == grid)]
    if not c:
        return None
    if which == "fastest":
        return min(c, key=lambda r: val(r, "gpu__time_duration.sum"))
    if which == "slowest":
        return max(c, key=lambda r: val(r, "gpu__time_duration.sum"))
    return c[-1]


LABELS = [
    ("spmm_csr_kernel (UI graph, fused layer-sum)", "spmm_csr_kernel<16, 1, 0", "fastest"),
    ("spmm_csr_kernel (scaled UI graph 600k x 120k, > L2)", "spmm_csr_kernel<16, 1, 0", "slowest"),
    ("gemm_tc05_kernel fwd (7050x4096 -> 64)", "gemm_tc05_kernel<0, 0, 64, 0, 0>", "last"),
    ("gemm_tc05_kernel dW", "gemm_tc05_kernel<1, 1, 64, 1, 0>", "last"),
    ("gemm_tc05_kernel<adam> (image table, low-rank gradient)", "gemm_tc05_kernel<0, 1, 32, 0, 1>", "last"),
    ("score_topk_tc_kernel (9130 x 7050, K=50)", "score_topk_tc_kernel", "last"),
    ("side_fwd_kernel (mma.sync)", "side_fwd_kernel", "last"), ("side_fwd_tc_kernel (tcgen05, shipped for d = 64)", "side_fwd_tc_kernel", "last"),
    ("side_bwd_kernel", "side_bwd_kernel", "last"),
    ("infonce_tc_kernel fwd", "infonce_tc_kernel<64, 0>", "last"),
    ("infonce_tc_kernel bwd row", "infonce_tc_kernel<64, 1>", "last"),
    ("infonce_tc_kernel bwd col", "infonce_tc_kernel<64, 2>", "last"),
    ("adam_kernel", "adam_kernel", "last"),
]
out = {"source": f"{sys.argv[1]} (ncu --set full --clock-control none, scripts/ncu_kernels.py; dram__bytes_read.sum + "
                 "dram__bytes_write.sum of the warm launch)", "dram_bytes_per_launch": {}, "l2_to_sm_bytes_per_launch": {},
       "duration_us": {}}
for label, sub, which in LABELS:
    r = pick(sub, which)
    if r is None:
        continue
    out["dram_bytes_per_launch"][label] = int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    out["duration_us"][label] = round(val(r, "gpu__time_duration.sum") * 1e6, 2)
    if "lts__t_sectors_srcunit_tex_op_read.sum" in hdr:
        out["l2_to_sm_bytes_per_launch"][label] = int(float(r[hdr.index("lts__t_sectors_srcunit_tex_op_read.sum")]) * 32)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
