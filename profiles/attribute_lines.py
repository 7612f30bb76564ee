#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample attribution for one kernel of an .ncu-rep.

usage: attribute_lines.py <report.ncu-rep> <kernel-regex> <mangled-prefix> [out.txt]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line markers of the in-tree
libmpcgpu.so (built with -lineinfo).  Run in the build container (no GPU needed)."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kre, mangled = sys.argv[1:4]
out = open(sys.argv[4], "w") if len(sys.argv) > 4 else sys.stdout
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "model-predictive-control-tuning_b200", "csrc")
tmp = tempfile.mkdtemp()
dis = None
objs = sorted(os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build")) if f.endswith(".o"))
for k, obj in enumerate(objs):   # one object (and cubin) per translation unit; the .so is linked from these
    d = os.path.join(tmp, str(k)); os.mkdir(d)
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
    for cubin in os.listdir(d):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout
        if ".text." + mangled in txt:
            dis = txt.split("\n")
    if dis:
        break
assert dis is not None, "kernel not found in libmpcgpu.so"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
# the export holds one block per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and re.search(kre, r[1])]
assert starts, "no launch matches " + kre
hdr = rows[starts[0] + 1]
data = []
for r in rows[starts[0] + 2:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        break
    data.append(r)
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + mangled)][0]
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
ai, as_ = collections.Counter(), collections.Counter()
for k, r in zip(seq, data):
    ai[k] += int(r[iI]); as_[k] += int(r[iS])
ti, ts = sum(ai.values()), sum(as_.values())
files = {}
def line(f, n):
    if f not in files:
        p = os.path.join(CSRC, f)
        files[f] = open(p).read().split("\n") if os.path.exists(p) else []
    return files[f][n - 1].strip()[:100] if n - 1 < len(files[f]) else ""
print(f"kernel {kre}: {ti} warp-instructions, {ts} stall samples, {len(data)} SASS instructions", file=out)
ops = collections.Counter()
for r in data:
    t = r[hdr.index("Source")].split()
    ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[iI])
print("opcode mix (% of executed warp-instructions): " + ", ".join(f"{k} {100*v/ti:.1f}" for k, v in ops.most_common(14)), file=out)
st = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
print("stall reasons (% of samples): " + ", ".join(f"{k[6:]} {100*v/ts:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v), file=out)
print("%inst %samples  file:line  source", file=out)
for k, v in ai.most_common(45):
    print(f"{100*v/ti:5.1f} {100*as_[k]/ts:5.1f}  {k[0]}:{k[1]}  {line(*k)}", file=out)
