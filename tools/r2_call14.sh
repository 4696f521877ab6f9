#!/bin/bash
# power: what do pure store streams draw when they run back to back, next to the skinning kernel?
mkdir -p gpurun_out
for pat in 0 4; do
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader -lms 100 > gpurun_out/c14_smi_$pat.txt &
SMI=$!
sleep 0.5
tools/ubench_store.bin sustained $pat 800 > gpurun_out/c14_store_$pat.txt 2>&1
kill $SMI
sleep 3
done
tail -n 4 gpurun_out/c14_store_0.txt gpurun_out/c14_store_4.txt
tail -n 5 gpurun_out/c14_smi_0.txt gpurun_out/c14_smi_4.txt
