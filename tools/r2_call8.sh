#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c8_build.log 2>&1
timeout 300 python tests/skin_ring_check.py > gpurun_out/c8_ring.log 2>&1; echo "exit $?" >> gpurun_out/c8_ring.log
for rep in 1 2; do
DMK_TOOL_CASES=ws_caller_order,ws_clustered,ws_clustered_streams,ws_clustered_s2s,ws_caller_order_s2s,ws_clustered_streams_s2s timeout 600 python tools/bench_configs.py 3s >> gpurun_out/c8_variants.jsonl 2>> gpurun_out/c8_variants.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c8_ws_clustered_s2s python tools/skin_only.py clustered s2s > gpurun_out/c8_ncu.log 2>&1
tail -n 2 gpurun_out/c8_ring.log
python - <<'P'
import json
for l in open('gpurun_out/c8_variants.jsonl'):
    d=json.loads(l); print({k:round(v['skin_ms'],3) for k,v in d.items() if isinstance(v,dict)})
P
