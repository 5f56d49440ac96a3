#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pcg or vmult or theta_sweep or smoother or cheby or tail or lanes or concurrent" ) > $O/r2_c7_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_c7_pytest.log
for th in 0.25 0.7; do
  AMGB_NO_ROW_SORT=1 timeout 300 python tools/run_one.py --m 200 --theta $th --mode full --repeat 2 --timers > $O/r2_c7_full_m200_th${th}_nosort.log 2>&1
  timeout 300 python tools/run_one.py --m 200 --theta $th --mode full --repeat 2 --timers > $O/r2_c7_full_m200_th${th}_sort.log 2>&1
  AMGB_NO_ROW_SORT=1 timeout 300 python tools/run_one.py --m 200 --theta $th --mode full --repeat 3 > $O/r2_c7_plain_m200_th${th}_nosort.log 2>&1
  timeout 300 python tools/run_one.py --m 200 --theta $th --mode full --repeat 3 > $O/r2_c7_plain_m200_th${th}_sort.log 2>&1
done
tail -n 4 $O/r2_c7_pytest.log
grep -H "^setup" $O/r2_c7_plain*.log | cut -d'|' -f2 | head -20
grep -H "^setup" $O/r2_c7_plain*.log | cut -c1-75
grep -H "prolong_l0 \|smooth  \|residual  \|aux  \|smooth_l0\|prolong  \|restrict  " $O/r2_c7_full*.log
