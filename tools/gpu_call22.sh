#!/bin/bash
# forecast_steps > 1 on the GPU (lfmq_chain_*) + the fp32 parity suite after the refactor
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_forecast_chain.py -m gpu -q -x --timeout 120 ) > gpurun_out/r02_c22_chain.log 2>&1; echo "chain rc=$?" > gpurun_out/r02_c22_rc.txt
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 ) > gpurun_out/r02_c22_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02_c22_rc.txt
cat gpurun_out/r02_c22_rc.txt; tail -n 30 gpurun_out/r02_c22_chain.log; tail -n 3 gpurun_out/r02_c22_parity.log
