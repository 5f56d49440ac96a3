#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2_c9_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2_c9_pytest_gpu.log
tail -n 8 $O/r2_c9_pytest_gpu.log
cd amg-ann_b200/host
./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_a.csv > ../../$O/r2_c9_datagen_t8.log 2>&1
AMGB_NO_SMALL_LEVELS=1 ./amgb_datagen --m 46 --systems 64 --threads 8 --seed 0 --device-assembly 1 --out /tmp/dg_b.csv > ../../$O/r2_c9_datagen_t8_nosmall.log 2>&1
./amgb_datagen --m 46 --systems 64 --threads 4 --seed 0 --device-assembly 1 --out /tmp/dg_c.csv > ../../$O/r2_c9_datagen_t4.log 2>&1
cd ../..
tail -n 1 $O/r2_c9_datagen_*.log
AMGB_TRACE=1 timeout 120 python tools/run_one.py --m 46 --mode full --repeat 3 2>&1 | grep -E "^setup|coarsen" | tail -8
