#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
cd amg-ann_b200/host
for t in 1 2 4 8; do
  ./amgb_datagen --m 46 --systems 16 --threads $t --seed 0 --device-assembly 1 --out /tmp/dg_$t.csv > ../../$O/r2_c8_datagen_t$t.log 2>&1
done
./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_8b.csv > ../../$O/r2_c8_datagen_t8_64.log 2>&1
AMGB_NO_AUTO_RESERVE=1 ./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_8c.csv > ../../$O/r2_c8_datagen_t8_64_noauto.log 2>&1
cd ../..
tail -n 2 $O/r2_c8_datagen_*.log
AMGB_TRACE=1 timeout 120 python tools/run_one.py --m 46 --mode full --repeat 3 > $O/r2_c8_trace_m46.log 2>&1
tail -n 60 $O/r2_c8_trace_m46.log
