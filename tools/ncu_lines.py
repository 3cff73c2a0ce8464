"""Per-source-line warp-instruction counts of one kernel: joins the SASS page of an .ncu-rep (executed counts, in
program order) with nvdisasm -g line info of the same function in the built object.
usage: python tools/ncu_lines.py file.ncu-rep sqcc_b200/lib/jk_fused.o 'jk_pipe_kernelILi4ELi4ELi1ELb1ELi8ELi2ELb1ELb1E' [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
sel = ["-s", os.environ["NCU_LAUNCH"], "-c", "1"] if os.environ.get("NCU_LAUNCH") else []   # one launch of a multi-launch report
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
ist = hdr.index("Warp Stall Sampling (All Samples)")
sass = [(r[isrc], int(r[ia] or 0), int(r[ist] or 0)) for r in rows[2:] if len(r) > ia]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
infn, line, recs = False, None, []
for ln in dis:
    if ln.startswith(".text."):
        infn = pat in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        recs.append((line, m.group(1)))
print(f"sass in report: {len(sass)}  instructions in object: {len(recs)}")
n = min(len(sass), len(recs))
bad = sum(1 for i in range(n) if sass[i][0].split()[0 if not sass[i][0].startswith('@') else 1].split('.')[0] not in recs[i][1])
print(f"opcode mismatches in join: {bad}")
byline, stall = collections.Counter(), collections.Counter()
tot = 0
for i in range(n):
    byline[recs[i][0]] += sass[i][1]
    stall[recs[i][0]] += sass[i][2]
    tot += sass[i][1]
stot = sum(stall.values()) or 1
srcs = {}
print(f"total warp instructions {tot}")
for (f, l), c in byline.most_common(top):
    if f not in srcs:
        for root in ("sqcc_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().splitlines()
                break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    print(f"{100 * c / tot:5.1f}% inst {100 * stall[(f, l)] / stot:5.1f}% stall  {f}:{l:<4d} {text[:110]}")
