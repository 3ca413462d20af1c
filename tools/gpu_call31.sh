#!/bin/bash
# cluster forward kernel: h_t published by one TMA store from a shared-memory staging tile (alt = per-thread STG + release fence)
mkdir -p gpurun_out
O=gpurun_out/r02_c31
timeout 90 python tools/time_steps.py --steps 5 --predict-batch 4096 > ${O}_probe.txt 2>&1; rc=$?
echo "probe rc=$rc" > ${O}_rc.txt
if [ $rc -ne 0 ]; then cat ${O}_rc.txt; tail -n 5 ${O}_probe.txt; exit 0; fi
( timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > ${O}_tests.log 2>&1; echo "tests rc=$?" >> ${O}_rc.txt
echo "== main (TMA publish)" > ${O}_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 65536 >> ${O}_time.txt 2>&1; done
echo "== alt (STG publish)" >> ${O}_time.txt
LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 65536 >> ${O}_time.txt 2>&1
LFMQ_TRACE_FWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> ${O}_ftrace.txt
cat ${O}_rc.txt; tail -n 2 ${O}_tests.log; grep -E "==|train|predict" ${O}_time.txt; sed -n 2,4p ${O}_ftrace.txt
