#!/usr/bin/env python
"""Static SASS footprint of a kernel, total and per CUDA source line (instruction-cache budget).

  python tools/sass_size.py solve_kernelI4Srbd [lib]"""
import collections, os, re, subprocess, sys, tempfile
kernel = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "srbd_horizon_b200", "csrc", "libsddp.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
cnt = collections.Counter(); cur = None; inside = False; inl = None
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kernel in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln):
        cnt[cur] += 16
tot = sum(cnt.values())
print(f"total {tot / 1024:.1f} KB   (per-phase footprint: tools/ncu_lines.py prints it next to the stall samples)")
print("top lines")
for key, n in cnt.most_common(25):
    print(f"{n / 1024:7.1f} KB  {key}")
