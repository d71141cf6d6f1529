#!/usr/bin/env python3
"""Per-source-line profile: joins `ncu --page source --csv` (SASS view, per-instruction counters) with `nvdisasm -g` line markers.

usage: ncu_line_profile.py <ncu_sass.csv> <nvdisasm.txt> <mangled-function-substring> [topN]
Both listings are in address order for the function, so the i-th instruction of one is the i-th of the other.
"""
import collections
import csv
import re
import sys

csv_path, dis_path, fn = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
lines, cur, infn = [], None, False
for ln in open(dis_path):
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        infn = (".text." in ln and fn in ln)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
        lines.append(cur)
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
    import glob
    if f not in src_cache:
        p = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src_cache[f] = open(p[0]).read().splitlines() if p else []
    s = src_cache[f]
    return s[l - 1].strip()[:110] if 0 < l <= len(s) else ""
for (key, v) in sorted(inst.items(), key=lambda kv: -kv[1])[:topn]:
    f, l = key if key else ("?", 0)
    print(f"{100 * v / ti:5.1f}% inst {100 * smp[key] / max(ts, 1):5.1f}% smp  {f}:{l:<5d} {src(f, l)}")
