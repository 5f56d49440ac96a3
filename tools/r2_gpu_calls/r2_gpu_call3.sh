#!/bin/bash
# flattened SpGEMM kernels under ncu: full capture of the level-0 pair, exported per instruction
set -u
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spgemm_flat -c 2 \
   -o $O/r2_c3_full_flat -f python tools/run_one.py --m 200 --mode setup > $O/r2_c3_b.log 2>&1
ncu -i $O/r2_c3_full_flat.ncu-rep --page raw --csv > $O/r2_c3_full_flat_raw.csv 2>/dev/null
ncu -i $O/r2_c3_full_flat.ncu-rep --page source --csv --print-kernel-base function > $O/r2_c3_full_flat_source.csv 2>/dev/null
gzip -f $O/r2_c3_full_flat_source.csv
rm -f $O/r2_c3_full_flat.ncu-rep
