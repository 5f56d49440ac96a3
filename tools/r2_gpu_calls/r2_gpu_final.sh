#!/bin/bash
# round-end check on one B200: full GPU test suite, smoke(), default bench line
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2_final_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2_final_pytest_gpu.log
tail -n 7 $O/r2_final_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2_final_smoke.log 2>&1
tail -n 5 $O/r2_final_smoke.log
( time timeout 1200 python bench.py ) > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err
echo "bench rc=$?" >> $O/r2_final_bench_n1.err
python - <<PY
import json
d=json.load(open("$O/r2_final_bench_n1.json"))
print("value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), d["clocks"], "wall", d.get("wall_s"), "traffic", d["roofline"]["traffic"], "frac", d["roofline"]["frac"])
print({k:(v.get("setup_ms"),v.get("solve_ms"),v.get("systems_per_s"),v.get("steady_systems_per_s"),v.get("value"),v.get("device_us"),v.get("images_per_s")) for k,v in d.get("configs",{}).items()})
PY
