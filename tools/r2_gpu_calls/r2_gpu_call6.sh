#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
run() { # name streams phases reserve
  BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 1 --e2e-steps 1 --extras none --streams $2 --phases $3 --reserve-gb $4 > $O/r2_c6_$1.json 2> $O/r2_c6_$1.err
  python - <<PY
import json
d=json.load(open("$O/r2_c6_$1.json"))
print("$1", "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), "clocks", d["clocks"]["sm_mhz"])
PY
  grep "bench\] sweep" $O/r2_c6_$1.err | tr '\n' ' '; echo
}
run s3p0r 3 0 20
run s3p0 3 0 0
run s2p0r 2 0 20
