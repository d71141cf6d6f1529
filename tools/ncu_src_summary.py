#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS lines. Usage: ncu_src_summary.py file.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[ix['# Samples']].isdigit()]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'warp instructions', sum(int(r[ix['Instructions Executed']]) for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f'{s:26s} {v:8d} {100 * v / max(tot, 1):5.1f}%')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:topn]:
    print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(10), r[ix['Source']][:100])
