#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "multicolour or substituted or smoother" ) > $O/r2_c18_pytest.log 2>&1
tail -n 25 $O/r2_c18_pytest.log | cut -c1-220
