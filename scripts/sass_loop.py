#!/usr/bin/env python
"""Static SASS summary of one kernel of core_b200/lib/mag_kernels.o: opcode histogram of the whole function and of
the address range between two PCs (the hot loop), to count instructions per entity before going to the GPU.
usage: sass_loop.py <substring of mangled name> [lo_pc hi_pc]"""
import collections, re, subprocess, sys
obj = "core_b200/lib/mag_kernels.o"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
args = [a for a in sys.argv[1:] if a != "-l"]
name = args[0]
lo = int(args[1], 16) if len(args) > 1 else 0
hi = int(args[2], 16) if len(args) > 2 else 1 << 30
cur, rows = None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and name in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
hist = collections.Counter()
for pc, ins in rows:
    if lo <= pc <= hi:
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        hist[op] += 1
n = sum(hist.values())
print("instructions in range: %d (function total %d)" % (n, len(rows)))
fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print("fp64-pipe: %d, other: %d" % (fp64, n - fp64))
print(", ".join("%s %d" % kv for kv in hist.most_common(30)))
if "-l" in sys.argv:
    for pc, ins in rows:
        if lo <= pc <= hi:
            print("%05x  %s" % (pc, ins))
