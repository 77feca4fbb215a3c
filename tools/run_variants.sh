#!/bin/bash
# A/B of kernel builds on the GPU box: tools/run_variants.sh DTYPE BATCH lib1.so lib2.so ...   (SDDP_LIB / SDDP_LIB_F32 overrides)
DT=$1; B=$2; shift 2
for L in "$@"; do
  if [ "$DT" = "f32" ]; then export SDDP_LIB_F32=$PWD/$L; else export SDDP_LIB=$PWD/$L; fi
  echo "== $L"; python tools/run_solve.py --batch $B --reps 4 --dtype $DT 2>&1 | tail -1
done
