"""Join an ncu SASS-level source page with nvdisasm line info and aggregate per source line / per function.

  ncu -i REPORT.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all libcfnmpc.so ; nvdisasm -gi -c X.cubin > dis_gi.txt
  python profiles/tools/ncu_by_line.py src.csv dis_gi.txt 'cf_rti_kernelILi4' [cf_rti_warp.h]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
focus = sys.argv[4] if len(sys.argv) > 4 else "cf_rti_warp.h"
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# ---- nvdisasm: offset -> (file,line) innermost location inside `focus`
loc = {}
in_k = False
chain = []
pending = []
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        in_k = kern in ln
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        if pending:
            chain = pending
            pending = []
        off = int(m.group(1), 16)
        pick = None
        for f, l in chain:
            if f.endswith(focus):
                pick = (f.split("/")[-1], l)
                break
        if pick is None and chain:
            pick = (chain[-1][0].split("/")[-1], chain[-1][1])
        loc[off] = pick

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
by_line = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_ex = tot_s = 0
ops = defaultdict(int)
for r in rows[2:]:
    off = int(r[ia], 16) - base
    ex, sm = int(r[iex] or 0), int(r[isamp] or 0)
    k = loc.get(off, ("?", 0))
    e = by_line[k]
    e[0] += ex
    e[1] += sm
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            e[2][hdr[i]] += v
    tot_ex += ex
    tot_s += sm
    op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    ops[op.split(".")[0]] += ex

print(f"total warp-instructions {tot_ex:,}  samples {tot_s:,}")
print("\n== top opcodes by executed count")
for op, v in sorted(ops.items(), key=lambda x: -x[1])[:25]:
    print(f"  {op:12s} {v:>16,} {100 * v / tot_ex:6.2f}%")

# function ranges of the focus file
funcs = []
try:
    path = [p for p in ("/root/repo/crazyflie_nmpc_b200/csrc/" + focus,) ][0]
    for n, l in enumerate(open(path), 1):
        m = re.match(r"\s*CF_(?:MEM|DEV)\s+[\w ]+?\s+\**(\w+)\(", l)
        if m:
            funcs.append((n, m.group(1)))
except Exception:
    pass


def func_of(line):
    name = "?"
    for n, f in funcs:
        if n <= line:
            name = f
    return name


by_func = defaultdict(lambda: [0, 0, defaultdict(int)])
for (f, l), (ex, sm, st) in by_line.items():
    k = func_of(l) if f == focus else f
    by_func[k][0] += ex
    by_func[k][1] += sm
    for a, b in st.items():
        by_func[k][2][a] += b
print("\n== by function: instr%  samples%  top stalls")
for k, (ex, sm, st) in sorted(by_func.items(), key=lambda x: -x[1][1]):
    top = ", ".join(f"{a[6:]}:{100 * b / max(sm, 1):.0f}%" for a, b in sorted(st.items(), key=lambda x: -x[1])[:4])
    print(f"  {k:26s} {100 * ex / tot_ex:6.2f}% {100 * sm / max(tot_s, 1):6.2f}%   {top}")
print(f"\n== top {topn} lines by samples")
for (f, l), (ex, sm, st) in sorted(by_line.items(), key=lambda x: -x[1][1])[:topn]:
    top = ", ".join(f"{a[6:]}:{100 * b / max(sm, 1):.0f}%" for a, b in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"  {f}:{l:<5d} instr {100 * ex / tot_ex:5.2f}%  samples {100 * sm / max(tot_s, 1):5.2f}%  {top}")
