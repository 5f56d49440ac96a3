#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_solve -c 1 \
    -o $O/r2_c16_dense -f python tools/run_one.py --m 200 --theta 0.95 --mode full --max-steps 2 > $O/r2_c16_ncu.log 2>&1
ncu -i $O/r2_c16_dense.ncu-rep --page raw --csv > $O/r2_c16_dense_raw.csv 2>/dev/null
ncu -i $O/r2_c16_dense.ncu-rep --page source --csv --print-kernel-base function > $O/r2_c16_dense_source.csv 2>/dev/null
gzip -f $O/r2_c16_dense_source.csv
rm -f $O/r2_c16_dense.ncu-rep
tail -n 2 $O/r2_c16_ncu.log
