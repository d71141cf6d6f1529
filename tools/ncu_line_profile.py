#!/usr/bin/env python3
"""Per-source-line profile: joins `ncu --page source --csv` (SASS view, per-instruction counters) with `nvdisasm -gi` line markers.

usage: ncu_line_profile.py <ncu_sass.csv> <nvdisasm.txt> <mangled-function-substring> [topN] [--outer | --depth D]
Both listings are in address order for the function, so the i-th instruction of one is the i-th of the other.
  default : attribute every SASS instruction to the innermost source line (the inlined leaf)
  --outer : attribute to the outermost frame of the inline chain (the line of the kernel body that led there); needs `nvdisasm -gi`
  --depth D : attribute to the D-th frame counted from the kernel body (0 = --outer)
"""
import collections
import csv
import glob
import re
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
csv_path, dis_path, fn = args[:3]
topn = int(args[3]) if len(args) > 3 else 25
depth = None
if "--outer" in sys.argv:
    depth = 0
if "--depth" in sys.argv:
    depth = int(sys.argv[sys.argv.index("--depth") + 1])
    args = [a for a in args if a != str(depth)] if len(args) > 4 else args

lines, chain, infn = [], [], False
pending = []
for ln in open(dis_path):
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        infn = (".text." in ln and fn in ln)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        pending.append(m)
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
        if pending:
            # a run of markers = one inline chain, innermost first; the last marker's "inlined at" is the outermost frame
            chain = [(pending[0].group(1).split("/")[-1], int(pending[0].group(2)))]
            for pm in pending:
                if pm.group(3):
                    chain.append((pm.group(3).split("/")[-1], int(pm.group(4))))
            pending = []
        if depth is None or not chain:
            lines.append(chain[0] if chain else None)
        else:
            frames = chain[::-1]   # outermost first
            lines.append(frames[min(depth, len(frames) - 1)])
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[ix["Instructions Executed"]].isdigit()]
print("sass in nvdisasm:", len(lines), " sass in ncu:", len(data))
n = min(len(lines), len(data))
inst, smp = collections.Counter(), collections.Counter()
for i in range(n):
    inst[lines[i]] += int(data[i][ix["Instructions Executed"]])
    smp[lines[i]] += int(data[i][ix["# Samples"]])
ti, ts = sum(inst.values()), sum(smp.values())
print("total warp instructions", ti, "samples", ts)
src_cache = {}


def src(f, l):
    if f not in src_cache:
        p = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src_cache[f] = open(p[0]).read().splitlines() if p else []
    s = src_cache[f]
    return s[l - 1].strip()[:110] if 0 < l <= len(s) else ""


for (key, v) in sorted(inst.items(), key=lambda kv: -kv[1])[:topn]:
    f, l = key if key else ("?", 0)
    print(f"{100 * v / ti:5.1f}% inst {100 * smp[key] / max(ts, 1):5.1f}% smp  {f}:{l:<5d} {src(f, l)}")
