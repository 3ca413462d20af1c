#!/bin/bash
# cluster backward kernel: partial exchange pushed through DSMEM (st.async) vs the L2 route (alt build)
mkdir -p gpurun_out
timeout 90 python tools/time_steps.py --steps 5 --predict-batch 4096 > gpurun_out/r02_c19_probe.txt 2>&1; rc=$?
echo "probe rc=$rc" > gpurun_out/r02_c19_rc.txt
if [ $rc -ne 0 ]; then cat gpurun_out/r02_c19_rc.txt; tail -n 5 gpurun_out/r02_c19_probe.txt; exit 0; fi
( timeout 200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > gpurun_out/r02_c19_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_c19_rc.txt
echo "== main (DSMEM push)" > gpurun_out/r02_c19_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c19_time.txt 2>&1; done
echo "== alt (L2 route)" >> gpurun_out/r02_c19_time.txt
for i in 1; do LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c19_time.txt 2>&1; done
LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c19_btrace.txt
LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c19_btrace_alt.txt
cat gpurun_out/r02_c19_rc.txt; tail -n 2 gpurun_out/r02_c19_tests.log; grep -E "==|train" gpurun_out/r02_c19_time.txt; head -5 gpurun_out/r02_c19_btrace.txt; head -5 gpurun_out/r02_c19_btrace_alt.txt
