#!/bin/bash
# timing experiment (results invalid in the alt build): general path without the row-major dz / h stores of the step epilogues
mkdir -p gpurun_out
O=gpurun_out/r02_c27
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_cfg3_main.json 2> ${O}_cfg3_main.err
LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_altA.so timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_cfg3_nostores.json 2> ${O}_cfg3_nostores.err
python - <<'PY'
import json
for n in ('main','nostores'):
    try:
        d=json.loads(open('gpurun_out/r02_c27_cfg3_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
