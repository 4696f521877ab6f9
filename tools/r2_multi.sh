#!/bin/bash
# round 2 multi-GPU call: both transports + the copy-engine one, content verified on the render GPU by bench.py itself
#   gpurun --gpus N --timeout 1200 -- 'bash tools/r2_multi.sh N'
set -u
N=${1:-2}
TAG=${2:-m}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}${N}_build.log 2>&1
PORT=29517
run() {  # name, extra bench args
  local name=$1; shift
  PORT=$((PORT + 1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N \
      --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}${N}_${name}.json 2> gpurun_out/${TAG}${N}_${name}.err
  echo "$name exit $?: $(python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${TAG}${N}_${name}.json').read().strip().splitlines()[-1])
    print(round(d['value']/1e6,2),'M chars/s', round(d['ms_per_step'],3),'ms/step e2e', round(d['e2e']['value']/1e6,2), 'verified', d.get('gather_verified'), 'skin', round(d['kernel_ms_per_step']['skin'],3), 'pack', round(d['kernel_ms_per_step']['pack_visible_palettes'],3), 'frame p90', round(d['frame_ms_rank0']['p90'],3))
except Exception as e:
    print('no line:', e)
P
)"
}
if [ "$N" = "8" ]; then
  run config4_peer --gather peer --chars 125000 --clips 32
  run config4_nccl --gather nccl --chars 125000 --clips 32
  run config4_peer_ce --gather peer-ce --chars 125000 --clips 32
  run peer --gather peer
  run peer_r0 --gather peer --reserve-sms 0
elif [ "$N" = "1" ]; then
  timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}1_bench.json 2> gpurun_out/${TAG}1_bench.err; tail -c 600 gpurun_out/${TAG}1_bench.json
else
  run peer_ce_r1 --gather peer-ce --reserve-sms 1
  run peer_ce_r0 --gather peer-ce --reserve-sms 0
  run peer_r4 --gather peer --reserve-sms 4
fi
tail -n 4 gpurun_out/${TAG}${N}_*.err | cut -c1-300
