"""Dynamic instruction count and stall samples per device function of one ncu report (needs the cubin for symbol sizes):
   python tools/ncu_by_function.py report.ncu-rep file.cubin"""
import csv, collections, subprocess, sys
rep, cubin = sys.argv[1], sys.argv[2]
out = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, ie, isamp, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = int(data[0][ia], 16)
exe = [(int(r[ia], 16) - base, int(r[ie]), int(r[isamp]), int(r[ith])) for r in data]
tot, tots = sum(e[1] for e in exe), sum(e[2] for e in exe)
sym = subprocess.run(f"readelf -sW {cubin}", shell=True, capture_output=True, text=True).stdout
funcs = []
for line in sym.splitlines():
    p = line.split()
    if len(p) >= 8 and p[3] == "FUNC":
        funcs.append((int(p[1], 16), int(p[2], 0), p[7]))
def fn(off):
    for a, sz, n in funcs:
        if n != "bo_solve_kernel" and a <= off < a + sz:
            return n.split("$")[-1]
    return "kernel body (master, fetch, write-out)"
agg, aggs, aggt = collections.Counter(), collections.Counter(), collections.Counter()
for off, e, s, t in exe:
    n = fn(off); agg[n] += e; aggs[n] += s; aggt[n] += t
print(f"total warp instructions {tot}  samples {tots}")
for n, v in agg.most_common():
    print(f"{n:48s} inst {v:12d} {100*v/tot:5.1f}%  stall samples {100*aggs[n]/tots:5.1f}%  avg threads {aggt[n]/max(1,v):5.1f}")
