#!/bin/bash
# final single-GPU evidence of the round: default bench line, lane-count A/B, ncu captures of the
# level-0 solve kernels and of the current setup kernels, pooling capture
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python bench.py ) > $O/r2_bench_n1_final.json 2> $O/r2_bench_n1_final.err
echo "bench rc=$?" >> $O/r2_bench_n1_final.err
for st in 2 4; do
  BENCH_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 1 --e2e-steps 1 --extras none --streams $st > $O/r2_c11_streams$st.json 2> $O/r2_c11_streams$st.err
done
export_rep() {  # report stem
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv --print-kernel-base function > $O/$1_source.csv 2>/dev/null
  gzip -f $O/$1_source.csv
  rm -f $O/$1.ncu-rep
}
AMGB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $O/r2_final_launches_full_m200.csv python tools/run_one.py --m 200 --mode full --max-steps 3 > $O/r2_c11_ncu_launches.log 2>&1
AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:sell_rows_kernel<1|sell_spmv_dot_kernel<1' -c 9 \
    -o $O/r2_final_full_solve_m200 -f python tools/run_one.py --m 200 --mode full --max-steps 2 > $O/r2_c11_ncu_solve.log 2>&1
export_rep r2_final_full_solve_m200
AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k 'regex:spgemm_flat|interp_fill_group' -c 3 \
    -o $O/r2_final_full_setup_m200 -f python tools/run_one.py --m 200 --mode setup > $O/r2_c11_ncu_setup.log 2>&1
export_rep r2_final_full_setup_m200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pool_entries -c 1 \
    -o $O/r2_final_full_pool_m200 -f python tools/run_one.py --m 200 --mode pool > $O/r2_c11_ncu_pool.log 2>&1
export_rep r2_final_full_pool_m200
python - <<PY
import json
for f in ("r2_bench_n1_final","r2_c11_streams2","r2_c11_streams4"):
    try:
        d=json.load(open("$O/"+f+".json")); print(f, "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), d["clocks"], d.get("wall_s"))
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 $O/r2_c11_ncu_solve.log
ls -la $O | head -40
