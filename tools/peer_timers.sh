#!/bin/bash
# per-family device time of the partitioned solve (event-bracketed launches) next to the wall time
N=${1:-2}; M=${2:-200}; TH=${3:-0.25}
run() {
  echo "== $1"
  shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 tools/dist_nccl.py --cells $M --theta $TH --repeat 3 --device-assembly $EXTRA 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|NCCL version" | tail -${TAIL:-3}
}
EXTRA=--timers TAIL=22 run "peer, no overlap, timers on the last repeat" AMGB_OVERLAP_MIN_ROWS=2000000000
EXTRA= run "peer, no overlap, graph" AMGB_OVERLAP_MIN_ROWS=2000000000 AMGB_DIST_GRAPH=1
EXTRA= run "nccl, graph" AMGB_PEER=0 AMGB_DIST_GRAPH=1
