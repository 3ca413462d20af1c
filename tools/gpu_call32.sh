#!/bin/bash
# general path: steps 1..T-1 of a layer as one persistent launch, barrier per row-tile group (LFMQ_GEN_PERSIST=0: one launch per step)
mkdir -p gpurun_out
O=gpurun_out/r02_c32
( timeout 100 python -m pytest tests/test_gpu_generic.py -m gpu -q -x --timeout 60 ) > ${O}_generic.log 2>&1; rc=$?
echo "generic rc=$rc" > ${O}_rc.txt
if [ $rc -ne 0 ]; then cat ${O}_rc.txt; tail -n 30 ${O}_generic.log; exit 0; fi
( timeout 200 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_bf16.py -m gpu -q -x --timeout 100 ) > ${O}_shapes.log 2>&1; echo "shapes rc=$?" >> ${O}_rc.txt
timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_cfg3_persist.json 2> ${O}_cfg3_persist.err; echo "cfg3 rc=$?" >> ${O}_rc.txt
LFMQ_GEN_PERSIST=0 timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_cfg3_stepped.json 2> ${O}_cfg3_stepped.err
cat ${O}_rc.txt; tail -n 3 ${O}_generic.log; tail -n 3 ${O}_shapes.log
python - <<'PY'
import json
for n in ('persist','stepped'):
    try:
        d=json.loads(open('gpurun_out/r02_c32_cfg3_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()}, 'launches', d['gpu_launches'])
    except Exception as e:
        print(n, 'ERR', e)
PY
