#!/bin/bash
# one GPU window of round-2 experiments (runs on the GPU box): correctness of the cell engine, engine and unfilter
# variants on the default bench, ncu captures.  Results in gpurun_out/c4_*.
mkdir -p gpurun_out
(timeout 240 python -m pytest tests/test_gpu_decode.py tests/test_gpu_zz_cells.py -q -x -k "cells or unfilter or pngsuite_all" 2>&1 | tail -4) > gpurun_out/c4_tests.log 2>&1
tools/gpu_exp.sh c4 main:6 serhdr:6 main:0 ub84:0 ub48:0 ub88:0 stage1:6 nostage:6 c2:6 > gpurun_out/c4_summary.txt 2>&1
cat gpurun_out/c4_tests.log gpurun_out/c4_summary.txt
# the whole GPU suite with the cell engine as the automatic choice (what flipping the default would ship), then as shipped
(PNGB200_CELLS_AUTO=1 timeout 500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4) > gpurun_out/c4_suite_cells_auto.log 2>&1
(timeout 500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4) > gpurun_out/c4_suite_default.log 2>&1
cat gpurun_out/c4_suite_cells_auto.log gpurun_out/c4_suite_default.log
N="python bench.py --workload 1080p-rgba8 --batch 444 --steps 1 --warmup 1 --no-e2e --no-cpu"
(timeout 240 ncu --set full --import-source on --clock-control none -k regex:inflate_cells -c 1 -f -o gpurun_out/c4_cells $N --inflate-mode 6 > gpurun_out/c4_ncu1.log 2>&1)
(timeout 200 ncu --set full --import-source on --clock-control none -k regex:unfilter_wave -c 1 -f -o gpurun_out/c4_unf_b1 $N > gpurun_out/c4_ncu2.log 2>&1)
(PNGB200_LIB=variants/libpngb200_ub84.so timeout 200 ncu --set full --import-source on --clock-control none -k regex:unfilter_wave -c 1 -f -o gpurun_out/c4_unf_b84 $N > gpurun_out/c4_ncu3.log 2>&1)
ls -la gpurun_out | tail -12
