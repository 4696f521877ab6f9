#!/bin/bash
# the driver's scaling command with the defaults of the final tree:  gpurun --gpus N -- 'bash tools/r2_scale.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s${N}_build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N \
    --steps 20 --warmup 5 > gpurun_out/s${N}_bench.json 2> gpurun_out/s${N}_bench.err
echo "exit $?"
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/s${N}_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']/1e6,2),'M chars/s', round(d['ms_per_step'],3),'ms/step e2e', round(d['e2e']['value']/1e6,2), 'verified', d.get('gather_verified'), d['kernel_ms_per_step'], d['frame_ms_rank0'], d['clocks'])
except Exception as e:
    print('no line:', e)
P
tail -n 5 gpurun_out/s${N}_bench.err | cut -c1-300
