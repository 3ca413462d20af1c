#!/bin/bash
# round 2, GPU call 3: chunked forward cluster kernel (tests + timing), ncu launch list + full captures of cfg3 step kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_c3_pytest.log 2>&1
timeout 300 python tools/time_steps.py --steps 20 > gpurun_out/r02_c3_time.txt 2>&1
LFMQ_TRACE_FWD=1 timeout 300 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> gpurun_out/r02_c3_ftrace.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 560 --csv --log-file gpurun_out/r02_c3_cfg3_launches.csv python tools/run_once.py --workload cfg3 --steps 2 > gpurun_out/r02_c3_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_gemm_kernel -s 261 -c 2 -o gpurun_out/r02_c3_fwdL2 python tools/run_once.py --workload cfg3 --steps 2 > gpurun_out/r02_c3_ncu2.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tile_gemm_kernel -s 309 -c 2 -o gpurun_out/r02_c3_bwd python tools/run_once.py --workload cfg3 --steps 2 > gpurun_out/r02_c3_ncu3.log 2>&1
tail -5 gpurun_out/r02_c3_pytest.log; cat gpurun_out/r02_c3_time.txt | grep train
