#!/bin/bash
# round 2, call 3: clustered path without zero fills: parity, timings, ncu
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c4_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_skin.py -q -m gpu -x > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c4_pytest.log
timeout 300 python tests/skin_ring_check.py > gpurun_out/c4_ring.log 2>&1; echo "exit $?" >> gpurun_out/c4_ring.log
DMK_TOOL_CASES=ws_caller_order,ws_sorted_by_influences,ws_clustered,ws_clustered_streams,ws_clustered_planes,round1_caller_order timeout 600 python tools/bench_configs.py 3s > gpurun_out/c4_variants.jsonl 2> gpurun_out/c4_variants.err
for shape in "clustered" "clustered planar3"; do
  name=$(echo "ws_${shape:-caller}" | tr ' ' '_')
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c4_${name} \
      python tools/skin_only.py $shape > gpurun_out/c4_ncu_${name}.log 2>&1
done
tail -n 3 gpurun_out/c4_pytest.log gpurun_out/c4_ring.log
cat gpurun_out/c4_variants.jsonl
