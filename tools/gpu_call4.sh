#!/bin/bash
# round 2, GPU call 4: general path with 256-row CTA tiles / deeper rings; cfg3 bench both ways; launch list
mkdir -p gpurun_out
( LFMQ_GEN_DUAL=1 timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -q -x ) > gpurun_out/r02_c4_generic_dual.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "generic or cfg3" ) > gpurun_out/r02_c4_generic.log 2>&1
LFMQ_GEN_DUAL=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_cfg3_dual0.json 2> gpurun_out/r02_c4_cfg3_dual0.err
LFMQ_GEN_DUAL=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_cfg3_dual1.json 2> gpurun_out/r02_c4_cfg3_dual1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_c4_cfg3_launches.csv python tools/run_once.py --workload cfg3 --steps 2 > gpurun_out/r02_c4_ncu1.log 2>&1
timeout 300 python bench.py --workload predict --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_predict_bf16.json 2> gpurun_out/r02_c4_predict_bf16.err
timeout 300 python bench.py --workload predict --precision bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_predict_x3.json 2> gpurun_out/r02_c4_predict_x3.err
tail -3 gpurun_out/r02_c4_generic_dual.log gpurun_out/r02_c4_generic.log
python - <<'PY'
import json
for f in ('r02_c4_cfg3_dual0','r02_c4_cfg3_dual1','r02_c4_predict_bf16','r02_c4_predict_x3'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), d['roofline']['regions_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
