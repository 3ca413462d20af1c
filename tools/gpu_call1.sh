#!/bin/bash
# round 2, GPU call 1: full gpu test suite, bench, backward trace, batcher bench, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_c1_smi.txt
( time python -m pytest tests -m gpu -q -k "not (cfg3_family and bf16)" ) > gpurun_out/r02_c1_pytest.log 2>&1
python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -s -k "not (cfg3_family and bf16)" > gpurun_out/r02_c1_shapes.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
LFMQ_TRACE_BWD=1 python tools/time_steps.py --steps 2 > gpurun_out/r02_c1_time.txt 2> gpurun_out/r02_c1_btrace.txt
python tools/time_steps.py --steps 20 >> gpurun_out/r02_c1_time.txt 2>&1
python bench.py --workload batcher --steps 50 > gpurun_out/r02_c1_batcher.json 2> gpurun_out/r02_c1_batcher.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_c1_ref.json 2>&1
tail -5 gpurun_out/r02_c1_pytest.log; cat gpurun_out/r02_c1_time.txt | tail -4
