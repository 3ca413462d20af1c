#!/bin/bash
# round 2, GPU call 2: general tensor-core path tests; backward recurrence with pipelined dz stores (+ prefetch variants)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -s ) > gpurun_out/r02_c2_generic.log 2>&1
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_bf16.py tests/test_gpu_cli.py -m gpu -q -s > gpurun_out/r02_c2_shapes.log 2>&1
for pf in 0 1 2 3; do
  echo "== LFMQ_BWD_PREFETCH=$pf" >> gpurun_out/r02_c2_time.txt
  LFMQ_BWD_PREFETCH=$pf timeout 300 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c2_time.txt 2>&1
done
LFMQ_TRACE_BWD=1 LFMQ_BWD_PREFETCH=0 timeout 300 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c2_btrace.txt
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c2_cfg3.json 2> gpurun_out/r02_c2_cfg3.err
tail -15 gpurun_out/r02_c2_generic.log; tail -5 gpurun_out/r02_c2_shapes.log; cat gpurun_out/r02_c2_time.txt | grep -v predict
