"""Shared-memory wavefronts and global L1 requests per source line of one kernel of an ncu capture.

  ncu -i REPORT.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all libcfnmpc.so ; nvdisasm -gi -c X.cubin > dis_gi.txt
  python profiles/tools/ncu_wavefronts_by_line.py src.csv dis_gi.txt 'cf_rti_kernelILi4ELi4ELi2ELb0' 'cf_rti_kernel<(int)4, (int)4, (int)2' [header]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern, kname = sys.argv[1:5]
focus = sys.argv[5] if len(sys.argv) > 5 else "cf_rti_warp.h"
loc, in_k, chain, pending = {}, False, [], []
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
            chain, pending = pending, []
        pick = None
        for f, l in chain:
            if f.endswith(focus):
                pick = l
                break
        loc[int(m.group(1), 16)] = pick if pick is not None else (chain[-1][1] if chain else 0)
rows = list(csv.reader(open(src_csv)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = [i for i in secs if kname in rows[i][1]][0]
e = min([i for i in secs if i > s] + [len(rows)])
hdr = rows[s + 1]
H = {h: i for i, h in enumerate(hdr)}
base = int(rows[s + 2][H["Address"]], 16)
by = defaultdict(lambda: defaultdict(lambda: [0, 0, 0]))
tot = [0, 0, 0]
for r in rows[s + 2:e]:
    if len(r) < len(hdr):
        continue
    t = r[H["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ex = int(r[H["Instructions Executed"]] or 0)
    w = int(r[H["L1 Wavefronts Shared"]] or 0)
    g = int(r[H["L1 Tag Requests Global"]] or 0)
    if not (w or g):
        continue
    a = by[loc.get(int(r[H["Address"]], 16) - base, 0)][op]
    a[0] += ex; a[1] += w; a[2] += g
    tot[0] += ex; tot[1] += w; tot[2] += g
print(f"total: {tot[0]:,} memory instructions, {tot[1]:,} shared wavefronts, {tot[2]:,} global L1 requests")
lines = sorted(by.items(), key=lambda kv: -sum(a[1] + a[2] for a in kv[1].values()))
srcl = open("/root/repo/crazyflie_nmpc_b200/csrc/" + focus).read().split("\n") if len(sys.argv) <= 6 else open(sys.argv[6]).read().split("\n")
for l, ops in lines[:int(sys.argv[7]) if len(sys.argv) > 7 else 60]:
    w = sum(a[1] for a in ops.values()); g = sum(a[2] for a in ops.values())
    d = ", ".join(f"{o} x{a[0] / 65536 / 1:.0f}" for o, a in sorted(ops.items(), key=lambda kv: -kv[1][1] - kv[1][2]))
    text = srcl[l - 1].strip()[:90] if 0 < l <= len(srcl) else ""
    print(f"{l:5d}  smem {100 * w / max(tot[1], 1):5.2f}%  glob {100 * g / max(tot[2], 1):5.2f}%  [{d}]  {text}")
