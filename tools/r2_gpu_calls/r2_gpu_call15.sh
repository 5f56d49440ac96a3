#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q ) > $O/r2_c15_pytest.log 2>&1
tail -n 3 $O/r2_c15_pytest.log
timeout 300 python tools/run_one.py --m 200 --theta 0.95 --mode full --repeat 2 --timers > $O/r2_c15_full_m200_th0.95.log 2>&1
grep "^setup\|coarse " $O/r2_c15_full_m200_th0.95.log | cut -c1-100
timeout 300 python tools/run_one.py --m 200 --theta 0.95 --mode full --repeat 3 > $O/r2_c15_plain_m200_th0.95.log 2>&1
grep "^setup" $O/r2_c15_plain_m200_th0.95.log | cut -d'|' -f2
