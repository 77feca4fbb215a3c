#!/bin/bash
# A/B of the three dense products of a backward-pass node: FP64 tensor cores (product build) against vector FP64 register
# tiles (build_ab/nodmma.so = -DSDDP_NO_DMMA=1).  gpurun --timeout 900 -- 'bash tools/ab_dmma.sh TAG'
TAG=${1:-ab}
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum"
for v in dmma nodmma; do
  if [ $v = dmma ]; then unset SDDP_LIB; else export SDDP_LIB=$PWD/build_ab/nodmma.so; fi
  echo "== $v"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  python tools/run_solve.py --batch 8192 --reps 4 | tail -1
  python tools/run_solve.py --batch 1 --N 20 --reps 6 | tail -1
  ncu --metrics $M --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv python tools/run_solve.py --batch 4736 --reps 2 2>&1 | grep -E '^"' | awk -F'","' 'NR>1 {print "   ", $(NF-2), $(NF-1), $NF}'
done > gpurun_out/${TAG}_ab_dmma.txt 2>&1
cat gpurun_out/${TAG}_ab_dmma.txt
