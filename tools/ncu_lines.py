#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction warp-stall samples by CUDA source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep solve_kernelI4Srbd [top]

ncu's CSV source page is SASS-only; the line table comes from `nvdisasm -g` on the cubin of the
in-tree libsddp.so (must be the same build that was profiled)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "srbd_horizon_b200", "csrc", "libsddp.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, inside = {}, None, False
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kernel in ln
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
wf_line = collections.Counter()
base = None
by_line = collections.Counter(); ex_line = collections.Counter()
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
stall_by_line = collections.defaultdict(collections.Counter)
total = 0
for r in rows[h + 1:]:
    if len(r) <= isamp:
        continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = a
    key = addr2line.get(a - base, (None, ""))[0]
    s = int(r[isamp] or 0)
    by_line[key] += s; total += s
    ex_line[key] += int(r[iex] or 0)
    if iwf is not None:
        wf_line[key] += int(float(r[iwf] or 0))
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            stall_by_line[key][hdr[i]] += v
src_cache = {}
def src(key):
    if key is None:
        return ""
    f, n = key
    if f not in src_cache:
        p = os.path.join(root, "srbd_horizon_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
print(f"total samples {total}")
for key, s in by_line.most_common(top):
    st = ", ".join(f"{k[6:]}={v}" for k, v in stall_by_line[key].most_common(3))
    print(f"{100.0 * s / total:5.1f}%  inst={ex_line[key]:>11}  {key}  [{st}]  {src(key)}")

# ---- optional phase buckets: python tools/ncu_lines.py rep kernel top buckets
if len(sys.argv) > 4:
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
    tot_b = collections.Counter(); ex_b = collections.Counter(); wf_b = collections.Counter(); st_b = collections.defaultdict(collections.Counter)
    for key, s in by_line.items():
        name = "other/unattributed"
        if key is not None:
            for nm, ranges in spec.items():
                if any(key[0] == f and lo <= key[1] <= hi for f, lo, hi in ranges):
                    name = nm
        tot_b[name] += s; ex_b[name] += ex_line[key]; wf_b[name] += wf_line[key]
        for k2, v in stall_by_line[key].items():
            st_b[name][k2] += v
    tot_ex = sum(ex_b.values())
    tot_wf = max(1, sum(wf_b.values()))
    print("\nphase buckets: samples%  inst%  smem-wavefronts%  top stalls")
    for nm, s in tot_b.most_common():
        st = ", ".join(f"{k[6:]}={100.0 * v / max(s, 1):.0f}%" for k, v in st_b[nm].most_common(3))
        print(f"{100.0 * s / total:5.1f}%  {100.0 * ex_b[nm] / tot_ex:5.1f}%  {100.0 * wf_b[nm] / tot_wf:5.1f}%  {nm:32s} [{st}]")
