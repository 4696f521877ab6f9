#!/bin/bash
# 4 vertices per thread (two lanes per 8-vertex block) against the 8-vertex shape: correctness on hardware, then times
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c11_build.log 2>&1
DMK_SKIN_VPT=4 timeout 600 python tests/skin_ring_check.py > gpurun_out/c11_ring_vpt4.log 2>&1; echo "ring vpt4 exit $?" >> gpurun_out/c11_ring_vpt4.log
DMK_SKIN_VPT=4 DMK_SKIN_VPT4_THREADS=480 timeout 600 python tests/skin_ring_check.py --quick > gpurun_out/c11_ring_vpt4w.log 2>&1; echo "ring vpt4 wide exit $?" >> gpurun_out/c11_ring_vpt4w.log
for rep in 1 2; do
for cfg in "8 352" "4 352" "4 480"; do
set -- $cfg
echo "vpt $1 threads $2" >> gpurun_out/c11_variants.jsonl
DMK_SKIN_VPT=$1 DMK_SKIN_VPT4_THREADS=$2 DMK_TOOL_CASES=ws_caller_order,ws_clustered timeout 600 python tools/bench_configs.py 3s >> gpurun_out/c11_variants.jsonl 2>> gpurun_out/c11_variants.err
done
done
tail -2 gpurun_out/c11_ring_vpt4.log gpurun_out/c11_ring_vpt4w.log
python - <<'P'
import json
for l in open('gpurun_out/c11_variants.jsonl'):
    if l.startswith('vpt'): print(l.strip()); continue
    d=json.loads(l); print({k:round(v['skin_ms'],3) for k,v in d.items() if isinstance(v,dict)})
P
