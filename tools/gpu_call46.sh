#!/bin/bash
# backward L2-prefetch helper lead (steps ahead) re-tuned for the faster backward kernel
mkdir -p gpurun_out
O=gpurun_out/r02_c46
echo "" > ${O}_time.txt
for lead in 2 0 1 3 4 2; do
  echo "== lead $lead" >> ${O}_time.txt
  LFMQ_BWD_PREFETCH=$lead timeout 120 python tools/time_steps.py --steps 20 >> ${O}_time.txt 2>&1
done
grep -E "==|train" ${O}_time.txt
