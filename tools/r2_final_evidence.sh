#!/bin/bash
# round 2, call 37: the final tree once on one GPU: tests, smoke, bench lines, launch list, ncu captures of the final skinning kernels
set -u
mkdir -p gpurun_out
rm -f gpurun_out/c37_*
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c37_build.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/c37_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c37_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c37_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c37_smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c37_bench_reference.json 2> gpurun_out/c37_bench_reference.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err
timeout 400 python bench.py --mesh-order caller --no-cpu-baseline > gpurun_out/c37_bench_caller.json 2> gpurun_out/c37_bench_caller.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c37_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/c37_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c37_skin_clustered python tools/skin_only.py clustered > gpurun_out/c37_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:skin_ -s 2 -c 1 -f -o gpurun_out/c37_skin_caller python tools/skin_only.py > gpurun_out/c37_ncu2.log 2>&1
tail -n 3 gpurun_out/c37_pytest.log gpurun_out/c37_smoke.log
for f in c37_bench c37_bench_caller c37_bench_reference; do echo "== $f"; tail -c 600 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err | cut -c1-200; done
ls -la gpurun_out/c37_*.ncu-rep
