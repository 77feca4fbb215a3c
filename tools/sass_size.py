#!/usr/bin/env python
"""Static SASS footprint of a kernel: total, per solver phase (same marker-delimited buckets as tools/ncu_lines.py)
and per CUDA source line.  The hot loop of solve_kernel is instruction-fetch bound (profiles/README.md), so this is
the budget to watch while editing; no GPU needed.

  python tools/sass_size.py solve_kernelI4Srbd8SmemSrbd [lib]"""
import collections, os, re, subprocess, sys, tempfile
kernel = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "srbd_horizon_b200", "csrc", "libsddp.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
cnt = collections.Counter(); cur = None; inside = False
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
print(f"total {tot / 1024:.1f} KB")

csrc = os.path.join(root, "srbd_horizon_b200", "csrc")
def mark(f, marker, nth=1):
    n = 0
    for i, l in enumerate(open(os.path.join(csrc, f)), 1):
        if marker in l:
            n += 1
            if n == nth:
                return i
    return None
B, Mo, So = "sddp_backward_srbd.cuh", "sddp_model.cuh", "sddp_solver.cuh"
try:
    marks = [("bwd load + top", mark(B, "__device__ int SmemSrbdT<MT, LAT>::backward(")), ("bwd c1 (Quu, gap)", mark(B, "// ---- c1:")),
             ("bwd d1 (warp-0 LDL^T, E)", mark(B, "// ---- d1:")), ("bwd c2 (T = V fx)", mark(B, "// ---- c2:")),
             ("bwd c3 (fx^T T, fu^T T)", mark(B, "// ---- c3:")), ("bwd e call", mark(B, "// ---- e:")),
             ("bwd h (Wn = Es B, DMMA)", mark(B, "// ---- h:")), ("bwd f,g (syrk + gains, DMMA)", mark(B, "// ---- f:")),
             ("bwd mu path + model", mark(B, "if (mu != 0.0) {   // regularised step")), ("", 10 ** 9)]
    spec = {name: [(B, lo, marks[i + 1][1] - 1)] for i, (name, lo) in enumerate(marks[:-1])}
    spec["bwd helpers (dmma, rcp, rows, contract)"] = [(B, 1, mark(B, "__device__ int SmemSrbdT<MT, LAT>::backward(") - 1)]
    spec["model: accel/xdot/cost lanes"] = [(Mo, mark(Mo, "SDDP_DEV static void accel_pre("), mark(Mo, "static void pack(") - 1)]
    spec["model: pack"] = [(Mo, mark(Mo, "static void pack("), mark(Mo, "SDDP_DEV static int zmap_x") - 1)]
    spec["model: expand (lx, lxx, lux terms)"] = [(Mo, mark(Mo, "SDDP_DEV static int zmap_x"), mark(Mo, "struct Lip {") - 1)]
    spec["model: m3 / inertia helpers"] = [(Mo, 1, mark(Mo, "SDDP_DEV static void accel_pre(") - 1)]
    spec["forward_wave"] = [(So, mark(So, "__device__ void forward_wave("), mark(So, "struct SolveArgs") - 1)]
    spec["solve_one control, packs, defects"] = [(So, mark(So, "struct SolveArgs"), 10 ** 9), (So, 1, mark(So, "__device__ void forward_wave(") - 1)]
    size_b = collections.Counter()
    for key, n in cnt.items():
        name = "other/unattributed"
        if key is not None:
            for nm, ranges in spec.items():
                if any(key[0] == f and lo <= key[1] <= hi for f, lo, hi in ranges):
                    name = nm
        size_b[name] += n
    print("per phase (static; tools/ncu_lines.py gives the executed-often part of a capture)")
    for nm, n in size_b.most_common():
        print(f"{n / 1024:7.1f} KB  {nm}")
except TypeError:
    print("(phase markers not found in the sources; per-line table only)")
print("top lines")
for key, n in cnt.most_common(25):
    print(f"{n / 1024:7.1f} KB  {key}")
