#!/bin/bash
# A/B of the solve-phase exchanges on N GPUs: library (NCCL) exchanges, peer windows without
# and with the interior/boundary overlap, and the captured V-cycle on top.
#   tools/peer_ab.sh <ngpus> <cells> [theta]
N=${1:-2}; M=${2:-200}; TH=${3:-0.25}
run() {
  echo "== $1"
  shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 tools/dist_nccl.py --cells $M --theta $TH --repeat 3 --device-assembly $CHECK 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
}
CHECK=--check run "nccl exchanges" AMGB_PEER=0
CHECK= run "peer windows, no overlap" AMGB_OVERLAP_MIN_ROWS=2000000000
CHECK=--check run "peer windows + overlap" AMGB_X=1
CHECK= run "peer windows + overlap + graph" AMGB_DIST_GRAPH=1
