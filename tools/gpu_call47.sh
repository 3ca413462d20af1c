#!/bin/bash
# 2 GPUs, final build: bench under torchrun, seeded DP-through-the-CLI check (fp32 and bf16), gloo-free
mkdir -p gpurun_out
O=gpurun_out/r02_c47
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err; echo "bench2 rc=$?" > ${O}_rc.txt
timeout 300 python tools/dp_cli_check.py --precision fp32 > ${O}_dpcli_fp32.log 2>&1; echo "dpcli fp32 rc=$?" >> ${O}_rc.txt
timeout 300 python tools/dp_cli_check.py --precision bf16 > ${O}_dpcli_bf16.log 2>&1; echo "dpcli bf16 rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 4 ${O}_dpcli_fp32.log; tail -n 4 ${O}_dpcli_bf16.log; tail -n 1 ${O}_bench_2gpu.json | cut -c1-300
