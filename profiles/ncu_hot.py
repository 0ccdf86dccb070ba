#!/usr/bin/env python
"""Hottest instructions (stall samples) of one kernel from an .ncu-rep:
   python tools/ncu_hot.py rep.ncu-rep <kernel-regex> [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
hdr = rows[hi]
isrc, ismp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seen, d = set(), []
for r in rows[hi + 1:]:
    if len(r) <= iex or r[0] in seen or not r[iex].isdigit():
        continue
    seen.add(r[0]); d.append(r)
tot = sum(int(r[ismp] or 0) for r in d)
print("total samples", tot)
order = sorted(range(len(d)), key=lambda i: -int(d[i][ismp] or 0))[:N]
for i in order:
    r = d[i]
    st = {hdr[c][6:]: int(r[c]) for c in cols if r[c] not in ('0', '')}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"#{i:5d} {100 * int(r[ismp] or 0) / tot:5.2f}% x{r[iex]:>7s} {r[isrc][:70]:70s} {top}")
