#!/bin/bash
# 2 GPUs: bench under torchrun (PDL + NCCL on one stream), DP through the CLI
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c14_bench_2gpu.json 2> gpurun_out/r02_c14_bench_2gpu.err; echo "bench2 rc=$?" > gpurun_out/r02_c14_rc.txt
timeout 300 python tools/dp_cli_check.py --precision fp32 > gpurun_out/r02_c14_dpcli_fp32.log 2>&1; echo "dpcli fp32 rc=$?" >> gpurun_out/r02_c14_rc.txt
timeout 300 python tools/dp_cli_check.py --precision bf16 > gpurun_out/r02_c14_dpcli_bf16.log 2>&1; echo "dpcli bf16 rc=$?" >> gpurun_out/r02_c14_rc.txt
cat gpurun_out/r02_c14_rc.txt; tail -n 4 gpurun_out/r02_c14_dpcli_fp32.log; tail -n 4 gpurun_out/r02_c14_dpcli_bf16.log
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_c14_bench_2gpu.json').read().strip().splitlines()[-1])
    print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('final_loss_mse'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/r02_c14_bench_2gpu.err').read()[-1500:])
PY
