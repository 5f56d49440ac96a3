#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) > $O/r2_c14_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_c14_pytest.log
tail -n 6 $O/r2_c14_pytest.log
for th in 0.25 0.7; do
  timeout 300 python tools/run_one.py --m 200 --theta $th --mode setup --repeat 3 --timers > $O/r2_c14_setup_m200_th$th.log 2>&1
  grep -H "interp \|interp  \|^setup" $O/r2_c14_setup_m200_th$th.log | cut -c1-200
done
