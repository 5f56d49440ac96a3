#!/bin/bash
# One gpurun call: GPU test suite, default bench line, per-family timers, ncu launch list and
# --set full captures of the top kernels (each ncu pass only after the plain run exited 0).
# gpurun_out/ must stay under 64 MiB or nothing comes back: reports are exported to CSV on the
# box and dropped when they are large.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2_box.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2_pytest_gpu.log
( time timeout 1200 python bench.py ) > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
echo "bench rc=$?" >> $O/r2_bench_n1.err
timeout 300 python tools/run_one.py --m 200 --mode full --repeat 2 --timers > $O/r2_timers_m200_theta0.25.log 2>&1
export_rep() {  # report stem
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  ncu -i $O/$1.ncu-rep --page source --csv --print-kernel-base function > $O/$1_source.csv 2>/dev/null
  gzip -f $O/$1_source.csv
  sz=$(stat -c %s $O/$1.ncu-rep)
  if [ "$sz" -gt 25000000 ]; then rm -f $O/$1.ncu-rep; fi
}
if AMGB_NO_GRAPH=1 timeout 300 python tools/run_one.py --m 200 --mode full --max-steps 3 > $O/r2_runone_plain.log 2>&1; then
  AMGB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file $O/r2_launches_full_m200.csv python tools/run_one.py --m 200 --mode full --max-steps 3 > $O/r2_ncu_launches.log 2>&1
  # level-0 setup kernels (first launches of each name)
  AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
      -k 'regex:spgemm_numeric|spgemm_symbolic|interp_fill|interp_ac|strength_kernel' -c 10 \
      -o $O/r2_full_setup_m200 -f python tools/run_one.py --m 200 --mode setup > $O/r2_ncu_full_setup.log 2>&1
  export_rep r2_full_setup_m200
  # level-0 solve kernels of the first PCG step
  AMGB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
      -k 'regex:sell_rows_kernel<1|sell_spmv_dot_kernel<1' -c 9 \
      -o $O/r2_full_solve_m200 -f python tools/run_one.py --m 200 --mode full --max-steps 2 > $O/r2_ncu_full_solve.log 2>&1
  export_rep r2_full_solve_m200
fi
if timeout 120 python tools/run_one.py --m 200 --mode pool > $O/r2_pool_plain.log 2>&1; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:pool_entries -c 1 \
      -o $O/r2_full_pool_m200 -f python tools/run_one.py --m 200 --mode pool > $O/r2_ncu_pool.log 2>&1
  export_rep r2_full_pool_m200
fi
while [ "$(du -sm $O | cut -f1)" -gt 55 ]; do
  big=$(ls -S $O | head -1); echo "dropping $big" >> $O/r2_dropped.txt; rm -f "$O/$big"
done
ls -la $O
