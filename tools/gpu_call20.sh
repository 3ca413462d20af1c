#!/bin/bash
# cluster backward kernel A/B: A = chunk-1 operands requested right after the export stores; B = per-warp peer signalling; C = both
mkdir -p gpurun_out
echo "" > gpurun_out/r02_c20_rc.txt
echo "== main" > gpurun_out/r02_c20_time.txt
timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c20_time.txt 2>&1
for v in A B C; do
  L=$PWD/lfm_quant_b200/_lfmq_alt$v.so
  LFMQ_LIB_PATH=$L timeout 60 python tools/time_steps.py --steps 5 --predict-batch 4096 > gpurun_out/r02_c20_probe$v.txt 2>&1; rc=$?
  echo "probe $v rc=$rc" >> gpurun_out/r02_c20_rc.txt
  if [ $rc -ne 0 ]; then continue; fi
  ( LFMQ_LIB_PATH=$L timeout 120 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x ) > gpurun_out/r02_c20_tests$v.log 2>&1; echo "tests $v rc=$?" >> gpurun_out/r02_c20_rc.txt
  echo "== alt $v" >> gpurun_out/r02_c20_time.txt
  for i in 1 2; do LFMQ_LIB_PATH=$L timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c20_time.txt 2>&1; done
  LFMQ_LIB_PATH=$L LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c20_btrace$v.txt
done
cat gpurun_out/r02_c20_rc.txt; grep -E "==|train" gpurun_out/r02_c20_time.txt; for v in A B C; do sed -n 2,3p gpurun_out/r02_c20_btrace$v.txt; done
