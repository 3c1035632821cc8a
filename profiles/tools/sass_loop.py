#!/usr/bin/env python
"""sass_loop.py <lib.so> <kernel-substring> [--dump]: dumps the SASS of one kernel and prints, for every backward branch, the
loop body's instruction mix (opcode histogram) -- the quick CPU-side check before GPU time is spent (B200_PROFILING.md)."""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for fn in funcs[1:]:
    name = fn.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in fn.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("== %s: %d instructions" % (name[:100], len(ins)))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            body = ins[j:i + 1]
            h = collections.Counter()
            for _, tt in body:
                op = tt.split()[1] if tt.startswith("@") else tt.split()[0]
                h[op.split(".")[0]] += 1
            print("loop 0x%x..0x%x: %d instr: %s" % (ins[j][0], a, len(body), ", ".join("%s %d" % kv for kv in h.most_common())))
    if "--dump" in sys.argv:
        for a, t in ins:
            print("%05x  %s" % (a, t))
