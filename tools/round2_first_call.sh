#!/bin/bash
# First GPU call of round 2: everything that was written after round 1's GPU budget ran out, verified and timed once.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/r02_*.  Every step has its own time limit; a failing step does not stop the rest.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -rxX > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
timeout 200 python tests/peer_exchange_check.py > gpurun_out/r02_peer_exchange_check.log 2>&1; echo "exit $?" >> gpurun_out/r02_peer_exchange_check.log
timeout 300 python tests/peer_exchange_ipc_check.py > gpurun_out/r02_peer_exchange_ipc_check.log 2>&1; echo "exit $?" >> gpurun_out/r02_peer_exchange_ipc_check.log
timeout 300 python tests/skin_planar_check.py > gpurun_out/r02_skin_planar_check.log 2>&1; echo "exit $?" >> gpurun_out/r02_skin_planar_check.log
timeout 300 python tests/skin_ring_check.py > gpurun_out/r02_skin_ring_check.log 2>&1; echo "exit $?" >> gpurun_out/r02_skin_ring_check.log
timeout 300 python tests/crowd_fuzz_check.py 1 200 > gpurun_out/r02_crowd_fuzz_check.log 2>&1; echo "exit $?" >> gpurun_out/r02_crowd_fuzz_check.log
timeout 300 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
for layout in 1 2 3 4; do timeout 200 python bench.py --skin-layout $layout --no-cpu-baseline > gpurun_out/r02_bench_layout$layout.json 2> gpurun_out/r02_bench_layout$layout.err; done
timeout 300 python tools/bench_configs.py 3s > gpurun_out/r02_skin_variants.jsonl 2> gpurun_out/r02_skin_variants.err
timeout 300 python tools/bench_pipelined.py > gpurun_out/r02_pipelined.json 2> gpurun_out/r02_pipelined.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_store.bin tools/ubench_store.cu && timeout 120 tools/ubench_store.bin > gpurun_out/r02_ubench_store.log 2>&1
# ncu: one full capture per skinning shape (only after the plain runs above), launch 6 = a steady-state skin launch
for shape in "" planar1 planar2 planar3 planar4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_kernel -s 2 -c 1 -f -o gpurun_out/r02_skin_${shape:-default} \
      python tools/skin_only.py $shape > gpurun_out/r02_ncu_skin_${shape:-default}.log 2>&1
done
tail -3 gpurun_out/r02_pytest_gpu.log gpurun_out/r02_peer_exchange_check.log gpurun_out/r02_peer_exchange_ipc_check.log gpurun_out/r02_skin_planar_check.log gpurun_out/r02_skin_ring_check.log gpurun_out/r02_crowd_fuzz_check.log gpurun_out/r02_skin_variants.jsonl
