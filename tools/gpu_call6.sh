#!/bin/bash
# round 2, GPU call 6: general path: tensor-core head, bwd epilogue with operand prefetch / 8 warps, faster aux kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_bf16.py -m gpu -q -x ) > gpurun_out/r02_c6_pytest.log 2>&1
( LFMQ_GEN_DUAL=1 timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -q -x ) > gpurun_out/r02_c6_generic_dual.log 2>&1
LFMQ_TRACE_GEN=1 timeout 300 python tools/run_once.py --workload cfg3 --steps 2 > /dev/null 2> gpurun_out/r02_c6_gtrace.txt
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c6_cfg3.json 2> gpurun_out/r02_c6_cfg3.err
LFMQ_GEN_DUAL=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c6_cfg3_dual0.json 2> gpurun_out/r02_c6_cfg3_dual0.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_c6_cfg3_launches.csv python tools/run_once.py --workload cfg3 --steps 2 > gpurun_out/r02_c6_ncu1.log 2>&1
timeout 300 python tools/time_steps.py --steps 20 > gpurun_out/r02_c6_time.txt 2>&1
tail -n 3 gpurun_out/r02_c6_pytest.log; tail -n 2 gpurun_out/r02_c6_generic_dual.log
python - <<'PY'
import json
for f in ('r02_c6_cfg3','r02_c6_cfg3_dual0'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -E "fwd l=1 t=(16|24)|bwd l=1 t=(16|24)" gpurun_out/r02_c6_gtrace.txt | head -8; grep train gpurun_out/r02_c6_time.txt
