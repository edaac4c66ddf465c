#!/bin/bash
# template of one GPU window (runs on the GPU box through tools/gpurun_retry.sh): the driver's checks on the shipped
# defaults and the records profiles/ keeps.  Edit per experiment; tools/gpu_exp.sh benchmarks tuning builds side by side.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4) > gpurun_out/suite.log 2>&1
cat gpurun_out/suite.log
(timeout 420 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err)
head -c 600 gpurun_out/bench_default.json; echo
