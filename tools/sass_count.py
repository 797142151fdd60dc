#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of an object file (cuobjdump -sass)."""
import collections, re, subprocess, sys
obj = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if m and fn:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[fn][op.split(".")[0]] += 1
for f, h in sorted(hist.items(), key=lambda kv: sum(kv[1].values())):
    if pat and pat not in f: continue
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(sum(h.values()), d, dict(h.most_common(14)))
