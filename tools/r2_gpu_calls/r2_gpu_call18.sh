#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for th in 0.25 0.7; do for sm in l1jacobi mcgs mcfb; do
  timeout 300 python tools/run_one.py --m 200 --theta $th --smoother $sm --mode full --repeat 2 > $O/r2_c18_m200_th${th}_$sm.log 2>&1
  echo "$th $sm: $(grep '^setup' $O/r2_c18_m200_th${th}_$sm.log | tail -1 | sed 's/levels.*opcx/opcx/')"
done; done
