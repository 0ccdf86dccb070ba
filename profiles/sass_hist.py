#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode from an .ncu-rep source page:
   python profiles/sass_hist.py rep.ncu-rep <kernel-regex> [pixels]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
px = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, iex, ith, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops, samples = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit():
        continue
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'ATOMS', 'BAR', 'SHFL', 'VOTE', 'F2I', 'I2F', 'FRND')) and '.' in op else '')
    ops[op] += int(r[iex]); samples[op] += int(r[ismp] or 0); tot += int(r[iex])
stot = sum(samples.values())
print(f"total warp instructions {tot}" + (f" = {tot*32/px:.1f} lane-instr/pixel" if px else ""))
for op, c in ops.most_common(40):
    print(f"{op:14s} {c:12d} {100*c/tot:5.1f}%   samples {100*samples[op]/max(stot,1):5.1f}%" + (f"   {c*32/px:6.2f}/px" if px else ""))
