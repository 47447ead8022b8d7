#!/usr/bin/env python
"""Executed-instruction mix (by SASS opcode) of one launch in an ncu report, beside the static
opcode histogram of the same kernel in the built library.
usage: sass_mix.py <report.ncu-rep> <launch index, -1 = last> <mangled kernel name> [lib.so]"""
import collections, csv, io, re, subprocess, sys

rep, which, mangled = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else "ldagibbssampling_b200/libb200lda.so"
OP = re.compile(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(raw)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif r and cur is not None and len(r) > 5:
        cur["rows"].append(r)
launches = kern[1::2] if len(kern) % 2 == 0 and kern[0]["name"] == kern[1]["name"] else kern  # ncu lists SASS and source views
k = launches[which]
h = k["hdr"]
ie, src = h.index("Instructions Executed"), h.index("Source")
dyn = collections.Counter()
for r in k["rows"]:
    m = OP.match(r[src])
    dyn[m.group(1) if m else "?"] += int(r[ie])
tot = sum(dyn.values())
sass = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, lib], capture_output=True, text=True).stdout
stat = collections.Counter()
for l in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", l)
    if m:
        stat[m.group(1)] += 1
print(f"kernel: {k['name']}")
print(f"executed warp instructions in this launch: {tot}; static SASS instructions: {sum(stat.values())}")
print(f"{'opcode':12s} {'executed %':>10s} {'static count':>13s}")
for op, n in dyn.most_common():
    if n * 1000 >= tot:
        print(f"{op:12s} {100 * n / tot:10.1f} {stat.get(op, 0):13d}")
