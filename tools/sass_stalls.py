#!/usr/bin/env python
"""Attribute the warp-stall samples of an ncu source-page CSV (tools/capture_source.sh) to source lines, using the
line table of the cubin (nvdisasm -g -c).  usage: sass_stalls.py <source.csv> <disassembly.txt> <function-substring> [top]"""
import collections
import csv
import re
import sys

src_csv, dis, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur_fn, chain, fresh = None, [], True
HELPER_MAX_LINE = 148  # cmul2 / cfnma2 / dmma884 ... : attribute their instructions to the caller (nvdisasm -gi chain)
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False  # nvdisasm prints the location only where it changes
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        if cur_fn and fn in cur_fn and chain:
            pick = next((c for c in chain if c[0] == "chan_kernels.cu" and c[1] > HELPER_MAX_LINE), chain[-1])
            line_of[int(m.group(1), 16)] = ((pick[0], pick[1], False), m.group(2))
        fresh = True
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp = hdr.index("Address"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
base = None
by_line = collections.Counter()
by_line_reason = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    n = int(r[isamp] or 0)
    tot += n
    ln = line_of.get(a - base, (None, "?"))[0]
    key = (ln[0], ln[1]) if ln else None
    by_line[key] += n
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            by_line_reason[key][hdr[i]] += v
print("total samples", tot)
for k, v in by_line.most_common(top):
    rs = ", ".join(f"{a[6:]} {100 * b / max(1, sum(by_line_reason[k].values())):.0f}%" for a, b in by_line_reason[k].most_common(3))
    print(f"{100 * v / tot:5.1f}%  {k}  [{rs}]")
