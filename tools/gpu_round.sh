#!/bin/bash
mkdir -p gpurun_out
tools/gpu_exp.sh c10 main:0 d16w4l2:0 l2b:0 l2c:0 > gpurun_out/c10_summary.txt 2>&1
cat gpurun_out/c10_summary.txt
