#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/e2e_probe.py 200 > $O/r2_c5_e2e_probe.log 2>&1
cd amg-ann_b200/host
for t in 8 16 32; do
  ( time ./amgb_datagen --m 46 --systems 64 --threads $t --seed 0 --device-assembly 1 --out /tmp/dg_$t.csv ) > ../../$O/r2_c5_datagen_t$t.log 2>&1
done
cd ../..
cat $O/r2_c5_e2e_probe.log; tail -n 8 $O/r2_c5_datagen_t*.log
