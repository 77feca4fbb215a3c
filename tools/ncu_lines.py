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
lib = os.environ.get("SDDP_NCU_LIB", os.path.join(root, "srbd_horizon_b200", "csrc", "libsddp.so"))      # SDDP_NCU_LIB: e.g. the fp32 build
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
ex_addr = {}
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
    ex_addr[a - base] = int(r[iex] or 0)
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
    # phase buckets are delimited by marker strings in the sources (first match), so they follow the code
    csrc = os.path.join(root, "srbd_horizon_b200", "csrc")
    def ln(f, marker, nth=1):
        n = 0
        for i, l in enumerate(open(os.path.join(csrc, f)), 1):
            if marker in l:
                n += 1
                if n == nth:
                    return i
        raise SystemExit(f"marker not found: {f}: {marker}")
    B, Mo, So = "sddp_backward_srbd.cuh", "sddp_model.cuh", "sddp_solver.cuh"
    marks = [("bwd load + top", ln(B, "__device__ int SmemSrbdT<MT, LAT>::backward(")), ("bwd c1 (Quu, gap)", ln(B, "// ---- c1:")),
             ("bwd d1 (warp-0 LDL^T, E)", ln(B, "// ---- d1:")), ("bwd c2 (T = V fx)", ln(B, "// ---- c2:")),
             ("bwd c3 (fx^T T, fu^T T)", ln(B, "// ---- c3:")), ("bwd e call", ln(B, "// ---- e:")),
             ("bwd h (Wn = Es B, DMMA)", ln(B, "// ---- h:")), ("bwd f,g (syrk + gains, DMMA)", ln(B, "// ---- f:")),
             ("bwd mu path + model", ln(B, "if (mu != 0.0) {")), ("", 10 ** 9)]
    spec = {name: [(B, lo, marks[i + 1][1] - 1)] for i, (name, lo) in enumerate(marks[:-1])}
    spec["bwd row helpers (axpy/store, in d1)"] = [(B, ln(B, "SDDP_DEV void axpy_row"), ln(B, "// out[b] = sum_a v[a] * (dt Aoo)") - 1)]
    spec["bwd dmma/rcp/contract helpers"] = [(B, ln(B, "SDDP_DEV void dmma884"), ln(B, "#ifndef SDDP_ROW128") - 1),
                                             (B, ln(B, "// out[b] = sum_a v[a] * (dt Aoo)"), ln(B, "__device__ int SmemSrbdT<MT, LAT>::backward(") - 1)]
    srbd0, lip0 = ln(Mo, "struct SrbdT {"), ln(Mo, "struct Lip {")
    spec["model: accel/xdot/cost lanes"] = [(Mo, ln(Mo, "SDDP_DEV static void accel("), ln(Mo, "static void pack(") - 1)]
    spec["model: pack (thread per node)"] = [(Mo, ln(Mo, "static void pack("), ln(Mo, "SDDP_DEV static int zmap_x") - 1)]
    spec["model: expand (lx, lxx, lux terms)"] = [(Mo, ln(Mo, "SDDP_DEV static int zmap_x"), lip0 - 1)]
    spec["model: m3 / inertia helpers"] = [(Mo, 1, ln(Mo, "SDDP_DEV static void accel(") - 1)]
    spec["forward_wave"] = [(So, ln(So, "__device__ void forward_wave("), ln(So, "struct SolveArgs") - 1)]
    spec["solve_one control, packs loop, defects"] = [(So, ln(So, "struct SolveArgs"), 10 ** 9), (So, 1, ln(So, "__device__ void forward_wave(") - 1)]
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
    size_b = collections.Counter(); hot_b = collections.Counter()
    mx = max(ex_addr.values()) if ex_addr else 1
    for a_, (key, _) in addr2line.items():
        name = "other/unattributed"
        if key is not None:
            for nm, ranges in spec.items():
                if any(key[0] == f and lo <= key[1] <= hi for f, lo, hi in ranges):
                    name = nm
        size_b[name] += 16
        if ex_addr.get(a_, 0) > 0.003 * mx:
            hot_b[name] += 16
    print(f"\nstatic SASS {sum(size_b.values()) / 1024:.0f} KB, executed often (> 0.3 % of the hottest instruction) {sum(hot_b.values()) / 1024:.0f} KB")
    print("\nphase buckets: samples%  inst%  smem-wavefronts%  SASS KB (hot KB)  top stalls")
    for nm, s in tot_b.most_common():
        st = ", ".join(f"{k[6:]}={100.0 * v / max(s, 1):.0f}%" for k, v in st_b[nm].most_common(3))
        print(f"{100.0 * s / total:5.1f}%  {100.0 * ex_b[nm] / tot_ex:5.1f}%  {100.0 * wf_b[nm] / tot_wf:5.1f}%  {size_b[nm] / 1024:5.1f} ({hot_b[nm] / 1024:4.1f})  {nm:38s} [{st}]")
