"""Dynamic opcode-class mix per source function: where the non-arithmetic instructions of a kernel are executed.

  ncu -i REPORT.ncu-rep --page source --csv > src.csv      (one kernel's SASS section)
  cuobjdump -xelf all libcfnmpc.so ; nvdisasm -gi -c X.cubin > dis_gi.txt
  python profiles/tools/ncu_opmix_by_function.py src.csv dis_gi.txt 'cf_rti_kernelILi4ELi4ELi2ELb0' [cf_rti_warp.h [SOURCE]]
(SOURCE: the header as it was when the binary was built, default the one in the tree -- function ranges come from it)

Classes: fp64 (DFMA DADD DMUL DMMA DSETP MUFU), smem (LDS STS), gmem (LDG STG LD ST ATOM RED), tma (UBLKCP SYNCS UTMA*),
shfl (SHFL), int (IMAD IADD3 VIADD LEA SHF LOP3 PRMT IABS ...), sel (FSEL SEL), pred (ISETP PLOP3 P2R R2P), mov (MOV
IMAD.MOV CS2R UMOV R2UR S2R ...), ctrl (BRA BSSY BSYNC NOP WARPSYNC EXIT ...)."""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
focus = sys.argv[4] if len(sys.argv) > 4 else "cf_rti_warp.h"
source = sys.argv[5] if len(sys.argv) > 5 else "/root/repo/crazyflie_nmpc_b200/csrc/" + focus

CLASSES = [("fp64", r"^(DFMA|DADD|DMUL|DMMA|DSETP|MUFU|F2F|I2F|F2I)"), ("smem", r"^(LDS|STS|LDSM)"),
           ("gmem", r"^(LDG|STG|LD|ST|ATOM|ATOMG|RED|LDC|LDCU|LDL|STL)\b"), ("tma", r"^(UBLKCP|SYNCS|UTMA|FENCE|MEMBAR|ERRBAR|CCTL)"),
           ("shfl", r"^(SHFL|VOTE|MATCH|REDUX)"), ("sel", r"^(FSEL|SEL|FMNMX|DMNMX)"),
           ("pred", r"^(ISETP|PLOP3|P2R|R2P|FSETP|ELECT)"), ("mov", r"^(MOV|IMAD\.MOV|CS2R|UMOV|R2UR|S2R|S2UR|PRMT|UIADD3|ULOP3|UISETP|USEL|ULEA|UIMAD|USHF)"),
           ("int", r"^(IMAD|IADD3|VIADD|LEA|SHF|LOP3|IABS|IMNMX|POPC|FLO|BREV|VIMNMX|IADD|ISCADD)"),
           ("ctrl", r"^(BRA|BSSY|BSYNC|NOP|WARPSYNC|EXIT|CALL|RET|BAR|YIELD|NANOSLEEP|BMOV|JMP|BRX)")]


def classify(op):
    for name, rx in CLASSES:
        if re.match(rx, op):
            return name
    return "other"


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
        loc[int(m.group(1), 16)] = pick

funcs = []
for n, l in enumerate(open(source), 1):
    m = re.match(r"\s*CF_(?:MEM|DEV)\s+[\w ]+?\s+\**(\w+)\(", l)
    if m:
        funcs.append((n, m.group(1)))


def func_of(line):
    name = "(kernel body)"
    if line is None:
        return name
    for n, f in funcs:
        if n <= line:
            name = f
    return name


rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
base = int(rows[2][ia], 16)
mix = defaultdict(lambda: defaultdict(int))
tot = 0
for r in rows[2:]:
    ex = int(r[iex] or 0)
    if not ex:
        continue
    s = r[isrc].split()
    op = s[1] if s[0].startswith("@") else s[0]
    mix[func_of(loc.get(int(r[ia], 16) - base))][classify(op)] += ex
    tot += ex
names = [c for c, _ in CLASSES] + ["other"]
print(f"total warp-instructions {tot:,}; per function: share of all instructions, then the function's own mix in %")
print(f"{'function':22s} {'all%':>6s} | " + " ".join(f"{n:>5s}" for n in names))
allmix = defaultdict(int)
for f, m in sorted(mix.items(), key=lambda kv: -sum(kv[1].values())):
    t = sum(m.values())
    for n in names:
        allmix[n] += m[n]
    if t < 0.002 * tot:
        continue
    print(f"{f:22s} {100 * t / tot:6.2f} | " + " ".join(f"{100 * m[n] / t:5.1f}" for n in names))
print(f"{'ALL':22s} {100.0:6.2f} | " + " ".join(f"{100 * allmix[n] / tot:5.1f}" for n in names))
