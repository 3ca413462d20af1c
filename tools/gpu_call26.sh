#!/bin/bash
# cluster backward kernel: exchange scratch in [16-column block][row][16] order (coalesced export stores, 1-D bulk loads); alt = row-major
mkdir -p gpurun_out
O=gpurun_out/r02_c26
timeout 90 python tools/time_steps.py --steps 5 --predict-batch 4096 > ${O}_probe.txt 2>&1; rc=$?
echo "probe rc=$rc" > ${O}_rc.txt
if [ $rc -ne 0 ]; then cat ${O}_rc.txt; tail -n 5 ${O}_probe.txt; exit 0; fi
( timeout 200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > ${O}_tests.log 2>&1; echo "tests rc=$?" >> ${O}_rc.txt
echo "== main (blocked exchange layout)" > ${O}_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> ${O}_time.txt 2>&1; done
echo "== alt (row-major exchange layout)" >> ${O}_time.txt
LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> ${O}_time.txt 2>&1
LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> ${O}_btrace.txt
cat ${O}_rc.txt; tail -n 2 ${O}_tests.log; grep -E "==|train" ${O}_time.txt; sed -n 2,4p ${O}_btrace.txt
