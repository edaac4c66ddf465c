#!/bin/bash
# one GPU window of round-2 experiments (runs on the GPU box).  Results in gpurun_out/c7_*.
mkdir -p gpurun_out
tools/gpu_exp.sh c7 ab:0 abp16:0 abp32:0 lvp16:0 > gpurun_out/c7_summary.txt 2>&1
cat gpurun_out/c7_summary.txt
(PNGB200_LIB=variants/libpngb200_abp16.so timeout 200 ncu --set full --import-source on --clock-control none -k regex:unfilter_wave -c 1 -f -o gpurun_out/c7_unf python bench.py --workload 1080p-rgba8 --batch 444 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/c7_ncu.log 2>&1)
