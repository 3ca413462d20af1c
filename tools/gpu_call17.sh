#!/bin/bash
# setmaxnreg in the cluster backward kernel (pointwise warpgroups 216 regs, role warpgroup 72): parity + timing; alt = LATE_C1 off
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > gpurun_out/r02_c17_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02_c17_rc.txt
( LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 200 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x ) > gpurun_out/r02_c17_tests_alt.log 2>&1; echo "alt tests rc=$?" >> gpurun_out/r02_c17_rc.txt
echo "== main (setmaxnreg 216/72, LATE_C1)" > gpurun_out/r02_c17_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c17_time.txt 2>&1; done
echo "== alt (setmaxnreg, LATE_C1=0)" >> gpurun_out/r02_c17_time.txt
for i in 1 2; do LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c17_time.txt 2>&1; done
cat gpurun_out/r02_c17_rc.txt; tail -n 2 gpurun_out/r02_c17_tests.log; tail -n 2 gpurun_out/r02_c17_tests_alt.log; grep -E "==|train" gpurun_out/r02_c17_time.txt
