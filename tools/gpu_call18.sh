#!/bin/bash
# setmaxnreg (one warpgroup-wide dec, 216/72) in the cluster backward kernel: quick probe first, then parity + timing
mkdir -p gpurun_out
timeout 90 python tools/time_steps.py --steps 5 --predict-batch 4096 > gpurun_out/r02_c18_probe.txt 2>&1; rc=$?
echo "probe rc=$rc" > gpurun_out/r02_c18_rc.txt
if [ $rc -ne 0 ]; then cat gpurun_out/r02_c18_rc.txt; tail -n 5 gpurun_out/r02_c18_probe.txt; exit 0; fi
( timeout 200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > gpurun_out/r02_c18_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_c18_rc.txt
( LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 200 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x ) > gpurun_out/r02_c18_tests_alt.log 2>&1; echo "alt tests rc=$?" >> gpurun_out/r02_c18_rc.txt
echo "== main (setmaxnreg 216/72, LATE_C1)" > gpurun_out/r02_c18_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c18_time.txt 2>&1; done
echo "== alt (setmaxnreg, LATE_C1=0)" >> gpurun_out/r02_c18_time.txt
for i in 1 2; do LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c18_time.txt 2>&1; done
cat gpurun_out/r02_c18_rc.txt; tail -n 2 gpurun_out/r02_c18_tests.log; tail -n 2 gpurun_out/r02_c18_tests_alt.log; grep -E "==|train" gpurun_out/r02_c18_time.txt
