#!/usr/bin/env python3
"""SASS instruction count of one kernel by source-line range (code size is a first-order cost of the closed-loop
kernel: the image has to stay inside the instruction cache).  usage: sass_size.py <object.o> <mangled-prefix> [bin]"""
import collections, os, re, subprocess, sys, tempfile
obj, mangled = os.path.abspath(sys.argv[1]), sys.argv[2]
binw = int(sys.argv[3]) if len(sys.argv) > 3 else 25
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
for cubin in os.listdir(d):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
    st = [i for i, l in enumerate(dis) if l.startswith(".text." + mangled)]
    if not st:
        continue
    cur, cnt = None, collections.Counter()
    for l in dis[st[0] + 1:]:
        if l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            cnt[cur] += 1
    print(mangled, sum(cnt.values()), "SASS instructions")
    rng = collections.Counter()
    for (f, n), v in cnt.items():
        rng[(f, n // binw * binw)] += v
    for k, v in sorted(rng.items()):
        print(f"  {k[0]}:{k[1]:<5d} {v}")
