#!/bin/bash
# round 2, call 2: warp-specialised skinning kernel + clustered meshes: parity, variants side by side, ncu of the winners
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c2_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_skin.py tests/test_gpu_fullsize.py tests/test_multi_skin_gpu.py -q -m gpu -x > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c2_pytest.log
timeout 300 python tests/skin_ring_check.py > gpurun_out/c2_ring.log 2>&1; echo "exit $?" >> gpurun_out/c2_ring.log
timeout 300 python tests/skin_planar_check.py > gpurun_out/c2_planar.log 2>&1; echo "exit $?" >> gpurun_out/c2_planar.log
timeout 300 python tests/crowd_fuzz_check.py 1 120 > gpurun_out/c2_fuzz.log 2>&1; echo "exit $?" >> gpurun_out/c2_fuzz.log
timeout 600 python tools/bench_configs.py 3s > gpurun_out/c2_variants.jsonl 2> gpurun_out/c2_variants.err
DMK_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
for shape in "" "clustered" "clustered planar3" "planar3"; do
  name=$(echo "ws_${shape:-caller}" | tr ' ' '_')
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c2_${name} \
      python tools/skin_only.py $shape > gpurun_out/c2_ncu_${name}.log 2>&1
done
tail -n 3 gpurun_out/c2_pytest.log gpurun_out/c2_ring.log gpurun_out/c2_planar.log gpurun_out/c2_fuzz.log
cat gpurun_out/c2_variants.jsonl
