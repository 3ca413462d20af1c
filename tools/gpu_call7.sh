#!/bin/bash
# round 2, GPU call 7: find the nondeterministic hang of call 6 (short timeouts)
mkdir -p gpurun_out
K='gradients or smaller or train_steps'
( LFMQ_DEBUG_SYNC=1 timeout 200 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "$K" ) > gpurun_out/r02_c7_sync.log 2>&1; echo "sync rc=$?" >> gpurun_out/r02_c7_rc.txt
( LFMQ_GEN_PDL=0 LFMQ_PDL=0 timeout 150 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "$K" ) > gpurun_out/r02_c7_nopdl.log 2>&1; echo "nopdl rc=$?" >> gpurun_out/r02_c7_rc.txt
( timeout 150 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "$K" ) > gpurun_out/r02_c7_default1.log 2>&1; echo "default1 rc=$?" >> gpurun_out/r02_c7_rc.txt
( timeout 150 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "$K" ) > gpurun_out/r02_c7_default2.log 2>&1; echo "default2 rc=$?" >> gpurun_out/r02_c7_rc.txt
( LFMQ_GEN_DUAL=1 timeout 150 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "$K" ) > gpurun_out/r02_c7_dual.log 2>&1; echo "dual rc=$?" >> gpurun_out/r02_c7_rc.txt
( timeout 200 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q -x ) > gpurun_out/r02_c7_cfg2.log 2>&1; echo "cfg2 rc=$?" >> gpurun_out/r02_c7_rc.txt
timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 > gpurun_out/r02_c7_time.txt 2>&1
LFMQ_PDL=0 timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c7_time.txt 2>&1
cat gpurun_out/r02_c7_rc.txt; tail -n 4 gpurun_out/r02_c7_sync.log | cut -c1-300; grep train gpurun_out/r02_c7_time.txt
