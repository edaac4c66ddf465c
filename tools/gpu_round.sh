#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -q -x -m gpu 2>&1 | tail -5) > gpurun_out/c12_suite.log 2>&1
cat gpurun_out/c12_suite.log
