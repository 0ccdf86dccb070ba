#!/usr/bin/env python
"""Executed-instruction mix per pixel of one kernel from an .ncu-rep (source page):
   python tools/ncu_mix.py rep.ncu-rep <kernel-regex> <pixels>"""
import collections, csv, io, re, subprocess, sys
rep, kern, px = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if 'Source' in r][0]
hdr = rows[hi]
isrc, ismp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
seen, ops, smp, tot = set(), collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) <= iex or r[0] in seen or not r[iex].isdigit():
        continue
    seen.add(r[0])
    n = int(r[iex])
    s = r[isrc].split()
    op = s[1] if s[0].startswith('@') else s[0]
    op = re.sub(r'^([A-Z0-9_]+(\.(128|64|RM|SAT|POPC))?).*', r'\1', op)
    ops[op] += n; tot += n; smp[op] += int(r[ismp] or 0)
px /= 32
stot = sum(smp.values())
print(f"total warp instr {tot} = {tot / px:.2f} lane-instr/px")
for op, n in ops.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    print(f"{op:12s} {n / px:6.2f}/px  samples {100 * smp[op] / stot:5.1f}%")
