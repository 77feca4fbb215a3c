#!/usr/bin/env python
"""Counts SASS opcodes of a kernel inside a source-line range (developer tool).
  python tools/sass_grep.py KERNEL FILE LO HI [lib]   e.g. solve_kernelI4Srbd8SmemSrbd sddp_backward_srbd.cuh 280 380"""
import collections, os, re, subprocess, sys, tempfile
kernel, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(root, "srbd_horizon_b200", "csrc", "libsddp.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
ops = collections.Counter(); inside = False; hit = False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kernel in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # innermost location or any frame of the inline chain inside the range
        locs = [(os.path.basename(m.group(1)), int(m.group(2)))] + [(os.path.basename(a), int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        hit = any(f == fname and lo <= n <= hi for f, n in locs)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and hit:
        ops[m.group(2).split(".")[0]] += 1
tot = sum(ops.values())
print(f"{tot} instructions ({tot * 16 / 1024:.1f} KB) in {fname}:{lo}-{hi} (inline chains included)")
print(", ".join(f"{k} {v}" for k, v in ops.most_common()))
