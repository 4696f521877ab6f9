#!/bin/bash
# round 2, call 17: the final tree once on one GPU: tests, smoke, bench lines (default, caller's order, reference), launch list
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c17_build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/c17_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c17_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c17_smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c17_bench_reference.json 2> gpurun_out/c17_bench_reference.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
timeout 400 python bench.py --mesh-order caller --no-cpu-baseline > gpurun_out/c17_bench_caller.json 2> gpurun_out/c17_bench_caller.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c17_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained > gpurun_out/c17_ncu_launches.log 2>&1
tail -n 3 gpurun_out/c17_pytest.log gpurun_out/c17_smoke.log
for f in c17_bench c17_bench_caller c17_bench_reference; do echo "== $f"; tail -c 900 gpurun_out/$f.json; echo; tail -n 2 gpurun_out/$f.err | cut -c1-200; done
