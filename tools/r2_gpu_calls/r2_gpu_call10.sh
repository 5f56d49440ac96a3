#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
cd amg-ann_b200/host
for i in 1 2; do
./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_a$i.csv > ../../$O/r2_c10_datagen_t8_$i.log 2>&1
AMGB_NO_SMALL_LEVELS=1 ./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_b$i.csv > ../../$O/r2_c10_datagen_t8_nosmall_$i.log 2>&1
done
cd ../..
tail -n 1 $O/r2_c10_datagen_*.log
