#!/bin/bash
# 2 GPUs: the real multi-GPU parity tests and the bench line with its `partitioned` block
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2_c13_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "distinct_gpus or nccl_processes" ) > $O/r2_c13_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> $O/r2_c13_pytest_2gpu.log
tail -n 6 $O/r2_c13_pytest_2gpu.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 ) > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
echo "bench rc=$?" >> $O/r2_bench_n2.err
tail -n 5 $O/r2_bench_n2.err
python - <<PY
import json
d=json.loads(open("$O/r2_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "wall", d.get("wall_s"))
print(json.dumps(d.get("partitioned"))[:1500])
PY
