#!/bin/bash
# Round-end measurement pass on one B200: every bench line the docs quote.   gpurun --timeout 1500 -- 'bash tools/final_bench.sh TAG'
TAG=${1:-fin}
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python bench.py --dtype f32 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_1gpu_f32.json 2> gpurun_out/${TAG}_bench_1gpu_f32.err
for c in 0 1 2 3; do python bench.py --config $c --steps 3 --warmup 3 --no-latency > gpurun_out/${TAG}_bench_config$c.json 2> gpurun_out/${TAG}_bench_config$c.err; done
python bench.py --dtype f32 --config 3 --steps 3 --warmup 3 --no-latency > gpurun_out/${TAG}_bench_config3_f32.json 2> gpurun_out/${TAG}_bench_config3_f32.err
python - $TAG <<'P'
import json,glob,sys
for f in sorted(glob.glob("gpurun_out/%s_*.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/fin_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], d.get("dtype"), "value %.1f e2e %.1f ms %.2f iters %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("mean_iters",0)), "cpu", d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("cores"), "parity", d.get("parity_max_rel_err"), "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
P
