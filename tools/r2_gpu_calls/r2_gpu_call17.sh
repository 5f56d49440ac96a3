#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tail or pcg or lanes or theta_sweep" ) > $O/r2_c17_pytest.log 2>&1
tail -n 3 $O/r2_c17_pytest.log
timeout 300 python tools/run_one.py --m 200 --theta 0.25 --mode full --repeat 2 --timers > $O/r2_c17_full_m200_th0.25.log 2>&1
grep "^setup\|tail " $O/r2_c17_full_m200_th0.25.log | cut -c1-120
timeout 300 python tools/run_one.py --m 100 --theta 0.25 --mode full --repeat 3 > $O/r2_c17_plain_m100.log 2>&1
grep "^setup" $O/r2_c17_plain_m100.log | cut -d'|' -f2
