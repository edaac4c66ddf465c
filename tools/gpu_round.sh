#!/bin/bash
# final GPU window of round 2: the driver's checks on the shipped defaults, the records for profiles/, then a last look
# at the L2-prefetch knob of the unfilter kernel.  Results in gpurun_out/c11_*.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4) > gpurun_out/c11_suite.log 2>&1
cat gpurun_out/c11_suite.log
(timeout 420 python bench.py > gpurun_out/c11_bench_default.json 2> gpurun_out/c11_bench_default.err)
head -c 600 gpurun_out/c11_bench_default.json; echo
(timeout 240 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum -k regex:inflate_parallel -c 1 --csv --log-file gpurun_out/c11_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1)
tail -4 gpurun_out/c11_traffic.csv
(timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/c11_launches.log 2>&1)
tools/gpu_exp.sh c11 d16w4l2:0 l2c:0 > gpurun_out/c11_summary.txt 2>&1
cat gpurun_out/c11_summary.txt
