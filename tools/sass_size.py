#!/usr/bin/env python
"""Static SASS footprint of a kernel per CUDA source line / phase bucket (instruction-cache budget).

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
spec = {
    "forward_wave": [("sddp_solver.cuh", 371, 502)],
    "node eval (accel/cost/xdot)": [("sddp_model.cuh", 193, 295), ("sddp_solver.cuh", 83, 107)],
    "init/defects/rollout": [("sddp_solver.cuh", 108, 172)],
    "pack (thread per node)": [("sddp_model.cuh", 296, 404)],
    "expand": [("sddp_model.cuh", 405, 677)],
    "bwd load+c1 (Quu, gap)": [("sddp_backward_srbd.cuh", 92, 184)],
    "bwd d1 (warp-0 LDL^T)": [("sddp_backward_srbd.cuh", 185, 239)],
    "bwd c2 (T=V fx)": [("sddp_backward_srbd.cuh", 240, 271)],
    "bwd c3 (Qxx,Qux cols)": [("sddp_backward_srbd.cuh", 272, 320)],
    "bwd d2 (RHS substitution)": [("sddp_backward_srbd.cuh", 321, 365)],
    "bwd syrk (DMMA)": [("sddp_backward_srbd.cuh", 366, 399)],
    "bwd K matmul (DMMA)": [("sddp_backward_srbd.cuh", 400, 427)],
    "bwd mu path + model": [("sddp_backward_srbd.cuh", 428, 490)],
    "solve_one control": [("sddp_solver.cuh", 503, 703)],
}
b = collections.Counter()
for key, n in cnt.items():
    name = "other"
    if key:
        for nm, rs in spec.items():
            if any(key[0] == f and lo <= key[1] <= hi for f, lo, hi in rs):
                name = nm
        if name == "other":
            name = "other:" + key[0]
    b[name] += n
tot = sum(b.values())
print(f"total {tot / 1024:.1f} KB")
for nm, n in b.most_common():
    print(f"{n / 1024:7.1f} KB  {nm}")
print("top lines")
for key, n in cnt.most_common(25):
    print(f"{n / 1024:7.1f} KB  {key}")
