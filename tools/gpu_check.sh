#!/bin/bash
# Developer loop on the GPU box: parity tests, then the two timing points used while tuning (B = 8192 and 65536).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG [quick]'
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_tests.log
python tools/run_solve.py --batch 8192 --reps 4 > gpurun_out/${TAG}_b8192.log 2>&1
if [ "$2" != "quick" ]; then
  python tools/run_solve.py --batch 65536 --reps 3 > gpurun_out/${TAG}_b65536.log 2>&1
  python tools/run_solve.py --batch 1 --reps 6 >> gpurun_out/${TAG}_b65536.log 2>&1
fi
cat gpurun_out/${TAG}_tests.log; grep -h "SUMMARY\|B=1 " gpurun_out/${TAG}_b8192.log gpurun_out/${TAG}_b65536.log 2>/dev/null | tail -8
