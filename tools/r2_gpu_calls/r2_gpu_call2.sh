#!/bin/bash
# flattened SpGEMM first stage: parity of the three variants, then A/B timing of the setup at m=200
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flattened or tiers or row_per_thread or aggressive or long_interp" ) > $O/r2_c2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_c2_pytest.log
for f in 0 1 2; do
  AMGB_SPGEMM_FLAT=$f timeout 300 python tools/run_one.py --m 200 --mode setup --repeat 3 --timers > $O/r2_c2_setup_m200_flat$f.log 2>&1
done
for f in 0 1; do
  AMGB_SPGEMM_FLAT=$f timeout 300 python tools/run_one.py --m 200 --theta 0.7 --mode setup --repeat 3 --timers > $O/r2_c2_setup_m200_th0.7_flat$f.log 2>&1
done
tail -n 3 $O/r2_c2_pytest.log
grep -h "spgemm" $O/r2_c2_setup_m200_*.log
