#!/bin/bash
# Profile pass on the GPU box: phase budget, launch list, one full ncu capture of solve_kernel with the per-line stall table.
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh TAG'
TAG=${1:-prof}
mkdir -p gpurun_out
if [ -f build_ab/libsddp_prof.so ]; then
  SDDP_LIB=$PWD/build_ab/libsddp_prof.so python tools/phase_timer.py --batch 4736 > gpurun_out/${TAG}_phases.txt 2>&1
fi
python tools/run_solve.py --batch 4736 --reps 3 > gpurun_out/${TAG}_b4736.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_solve \
    python tools/run_solve.py --batch 4736 --reps 2 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_solve.ncu-rep --page raw --csv > gpurun_out/${TAG}_solve_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_solve.ncu-rep --page details > gpurun_out/${TAG}_solve_details.txt 2>/dev/null
python tools/ncu_lines.py gpurun_out/${TAG}_solve.ncu-rep solve_kernelI5SrbdTILb0EE9SmemSrbdTIS1_Lb0EELi4EE 60 buckets > gpurun_out/${TAG}_solve_lines.txt 2>&1
tail -3 gpurun_out/${TAG}_b4736.log; head -40 gpurun_out/${TAG}_phases.txt
