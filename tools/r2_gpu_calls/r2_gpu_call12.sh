#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:sell_rows_kernel<\(int\)1, \(bool\)[01], amgb::EpiJacobi' -c 10 \
    -o $O/r2_final_full_smooth_m200 -f python tools/run_one.py --m 200 --mode full --max-steps 2 > $O/r2_c12_ncu_smooth.log 2>&1
ncu -i $O/r2_final_full_smooth_m200.ncu-rep --page raw --csv > $O/r2_final_full_smooth_m200_raw.csv 2>/dev/null
rm -f $O/r2_final_full_smooth_m200.ncu-rep
tail -n 3 $O/r2_c12_ncu_smooth.log
