#!/bin/bash
# Profile pass of the fp32 build on the GPU box (same steps as gpu_prof.sh).   gpurun --timeout 900 -- 'bash tools/gpu_prof_f32.sh TAG'
TAG=${1:-prof32}
mkdir -p gpurun_out
if [ -f build_ab/libsddp_f32_prof.so ]; then
  SDDP_LIB_F32=$PWD/build_ab/libsddp_f32_prof.so python tools/phase_timer.py --batch 5920 --dtype f32 > gpurun_out/${TAG}_phases.txt 2>&1
fi
python tools/run_solve.py --batch 5920 --reps 3 --dtype f32 > gpurun_out/${TAG}_b5920.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_solve \
    python tools/run_solve.py --batch 5920 --reps 2 --dtype f32 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_solve.ncu-rep --page raw --csv > gpurun_out/${TAG}_solve_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_solve.ncu-rep --page details > gpurun_out/${TAG}_solve_details.txt 2>/dev/null
SDDP_NCU_LIB=$PWD/srbd_horizon_b200/csrc/libsddp_f32.so python tools/ncu_lines.py gpurun_out/${TAG}_solve.ncu-rep solve_kernelI5SrbdTILb0EE9SmemSrbdTIS1_Lb0EELi5EE 60 buckets > gpurun_out/${TAG}_solve_lines.txt 2>&1
tail -3 gpurun_out/${TAG}_b5920.log; head -40 gpurun_out/${TAG}_phases.txt
