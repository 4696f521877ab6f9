#!/bin/bash
# Multi-GPU call of round 2 (N = 2 first, then 8): the NCCL gather after the equal-capacity fix, and the peer-memory exchange.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_multi_gpu.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
PORT=29517
run() {  # name, extra bench args
  local name=$1; shift
  PORT=$((PORT + 1))   # a fresh rendezvous port per run
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N \
      --steps 20 --warmup 5 "$@" > gpurun_out/r02_b${N}_${name}.json 2> gpurun_out/r02_b${N}_${name}.err
  echo "$name exit $?: $(cut -c1-200 gpurun_out/r02_b${N}_${name}.json)"
}
run nccl --gather nccl
run peer --gather peer
run peer_instream --gather peer --reserve-sms 0
if [ "$N" = "8" ]; then run config4_peer --gather peer --chars 125000 --clips 32; run config4_nccl --gather nccl --chars 125000 --clips 32; fi
