#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c5_build.log 2>&1
for rep in 1 2 3; do
DMK_TOOL_CASES=ws_caller_order,ws_caller_order_streams,ws_clustered,ws_clustered_streams timeout 600 python tools/bench_configs.py 3s >> gpurun_out/c5_variants.jsonl 2>> gpurun_out/c5_variants.err
done
python - <<'P'
import json
for l in open('gpurun_out/c5_variants.jsonl'):
    d=json.loads(l); print({k:round(v['skin_ms'],3) for k,v in d.items() if isinstance(v,dict)})
P
