#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c7_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_skin.py -q -m gpu -x > gpurun_out/c7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c7_pytest.log
timeout 300 python tests/skin_ring_check.py > gpurun_out/c7_ring.log 2>&1; echo "exit $?" >> gpurun_out/c7_ring.log
for rep in 1 2; do
DMK_TOOL_CASES=ws_caller_order,ws_clustered,ws_clustered_streams timeout 600 python tools/bench_configs.py 3s >> gpurun_out/c7_variants.jsonl 2>> gpurun_out/c7_variants.err
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c7_ws_clustered python tools/skin_only.py clustered > gpurun_out/c7_ncu.log 2>&1
tail -n 2 gpurun_out/c7_pytest.log gpurun_out/c7_ring.log
python - <<'P'
import json
for l in open('gpurun_out/c7_variants.jsonl'):
    d=json.loads(l); print({k:round(v['skin_ms'],3) for k,v in d.items() if isinstance(v,dict)})
d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), 'frac', round(d['roofline']['frac'],4), d['kernel_ms_per_step'])
P
