#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "not cfg3" ) > gpurun_out/r02_c15_cfg2.log 2>&1; echo "cfg2 rc=$?" > gpurun_out/r02_c15_rc.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c15_time.txt 2>&1; done
LFMQ_TRACE_BWD=1 LFMQ_BWD_PREFETCH=0 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c15_btrace.txt
cat gpurun_out/r02_c15_rc.txt; tail -n 2 gpurun_out/r02_c15_cfg2.log; grep train gpurun_out/r02_c15_time.txt; head -4 gpurun_out/r02_c15_btrace.txt
