#!/bin/bash
# SpGEMM first-stage variants: parity, then timing of the setup at m=200
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flattened or tiers or aggressive or long_interp or row_per_thread" ) > $O/r2_c4_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_c4_pytest.log
for th in 0.25 0.7; do
for f in 0 1 4 auto; do
  if [ $f = auto ]; then unset AMGB_SPGEMM_FLAT; else export AMGB_SPGEMM_FLAT=$f; fi
  timeout 300 python tools/run_one.py --m 200 --theta $th --mode setup --repeat 3 --timers > $O/r2_c4_setup_m200_th${th}_flat$f.log 2>&1
done
done
unset AMGB_SPGEMM_FLAT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv -k regex:spgemm \
   --log-file $O/r2_c4_launches_flat.csv python tools/run_one.py --m 200 --mode setup > $O/r2_c4_a.log 2>&1
tail -n 5 $O/r2_c4_pytest.log
grep -H "spgemm " $O/r2_c4_setup_m200_*.log
grep -H "^setup" $O/r2_c4_setup_m200_*.log | cut -c1-120
