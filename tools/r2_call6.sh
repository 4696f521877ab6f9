#!/bin/bash
# round 2, call 6: fused pose kernel: parity (whole -m gpu suite), pose A/B, bench, ncu of the fused kernel + launch list
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c6_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c6_pytest.log
for v in "" "pal44" "two pal44" "two"; do echo "== pose_only $v" >> gpurun_out/c6_pose.log; timeout 120 python tools/pose_only.py 50 $v >> gpurun_out/c6_pose.log 2>&1; done
# the other occupancy of the fused kernel
sed -i 's/#define DMK_FUSED_MIN_BLOCKS 3/#define DMK_FUSED_MIN_BLOCKS 2/' darmok_b200/csrc/pose_kernel.cu
python -c "import __graft_entry__ as g; g.build()" >> gpurun_out/c6_build.log 2>&1
echo "== pose_only (min blocks 2)" >> gpurun_out/c6_pose.log; timeout 120 python tools/pose_only.py 50 >> gpurun_out/c6_pose.log 2>&1
sed -i 's/#define DMK_FUSED_MIN_BLOCKS 2/#define DMK_FUSED_MIN_BLOCKS 3/' darmok_b200/csrc/pose_kernel.cu
python -c "import __graft_entry__ as g; g.build()" >> gpurun_out/c6_build.log 2>&1
timeout 300 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
timeout 300 python bench.py --mesh-order caller --no-cpu-baseline > gpurun_out/c6_bench_caller.json 2> gpurun_out/c6_bench_caller.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pose_fused -s 3 -c 1 -f -o gpurun_out/c6_pose_fused python tools/pose_only.py 5 > gpurun_out/c6_ncu_pose.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c6_ncu_launches.log 2>&1
tail -n 3 gpurun_out/c6_pytest.log; cat gpurun_out/c6_pose.log; tail -c 1500 gpurun_out/c6_bench.json
