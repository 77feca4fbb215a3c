M="gcc__cache_requests_type_instruction.sum,gpu__time_duration.sum,smsp__inst_executed.sum"
for v in "$@"; do
  echo "== $v"
  SDDP_LIB=$PWD/build_ab/$v.so python tools/run_solve.py --batch 8192 --reps 4 | tail -1
  SDDP_LIB=$PWD/build_ab/$v.so python tools/run_solve.py --batch 1 --reps 5 | tail -1
  SDDP_LIB=$PWD/build_ab/$v.so ncu --metrics $M --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv python tools/run_solve.py --batch 4736 --reps 2 2>&1 | grep -E "gcc__|gpu__time|smsp__inst" | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
