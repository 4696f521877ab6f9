#!/bin/bash
# round 2, call 9: everything once on one GPU with the final defaults
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c9_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c9_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c9_smoke.log
timeout 400 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
timeout 400 python bench.py --mesh-order caller --no-cpu-baseline > gpurun_out/c9_bench_caller.json 2> gpurun_out/c9_bench_caller.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c9_bench_reference.json 2> gpurun_out/c9_bench_reference.err
timeout 300 python bench.py --config 2 > gpurun_out/c9_bench_config2.json 2> gpurun_out/c9_bench_config2.err
timeout 400 python bench.py --config 5 > gpurun_out/c9_bench_config5.json 2> gpurun_out/c9_bench_config5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c9_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c9_skin_clustered python tools/skin_only.py clustered > gpurun_out/c9_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c9_skin_caller python tools/skin_only.py > gpurun_out/c9_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cull_kernel -s 3 -c 1 -f -o gpurun_out/c9_cull_10m python bench.py --config 5 --steps 5 --warmup 5 > gpurun_out/c9_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 3 -c 1 -f -o gpurun_out/c9_sample python tools/pose_only.py 5 > gpurun_out/c9_ncu4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:model_kernel -s 3 -c 1 -f -o gpurun_out/c9_model python tools/pose_only.py 5 > gpurun_out/c9_ncu5.log 2>&1
tail -n 3 gpurun_out/c9_pytest.log gpurun_out/c9_smoke.log
for f in c9_bench c9_bench_caller c9_bench_reference c9_bench_config2 c9_bench_config5; do echo "== $f"; tail -c 700 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err | cut -c1-200; done
