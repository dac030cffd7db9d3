#!/usr/bin/env python
"""Static SASS census of one kernel of an object file: instruction count and opcode mix (offline proxy for the
issue-bound kernels).  usage: tools/sass_count.py <object> <mangled-name substring>"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, ops = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name and pat in name:
        ops[name][m.group(1).split(".")[0]] += 1
for n, c in ops.items():
    tot = sum(c.values())
    print(n, tot)
    print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
