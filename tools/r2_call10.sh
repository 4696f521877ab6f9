#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c10_build.log 2>&1
for pf in 0 8 16 0 8 32; do
echo "prefetch $pf" >> gpurun_out/c10_variants.jsonl
DMK_SKIN_L2_PREFETCH=$pf DMK_TOOL_CASES=ws_caller_order,ws_clustered timeout 600 python tools/bench_configs.py 3s >> gpurun_out/c10_variants.jsonl 2>> gpurun_out/c10_variants.err
done
python - <<'P'
import json
for l in open('gpurun_out/c10_variants.jsonl'):
    if l.startswith('prefetch'): print(l.strip()); continue
    d=json.loads(l); print({k:round(v['skin_ms'],3) for k,v in d.items() if isinstance(v,dict)})
P
