#!/bin/bash
# full GPU suite with the persistent general-path steps and the cluster-path changes since c23; cfg3 / cfg2 bench lines
mkdir -p gpurun_out
O=gpurun_out/r02_c35
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > ${O}_pytest.log 2>&1; echo "pytest rc=$?" > ${O}_rc.txt
timeout 600 python bench.py > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err; echo "bench cfg2 rc=$?" >> ${O}_rc.txt
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > ${O}_bench_cfg3.json 2> ${O}_bench_cfg3.err; echo "bench cfg3 rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload predict --precision bf16x3 > ${O}_bench_predict_bf16x3.json 2> ${O}_bench_predict_bf16x3.err; echo "predict x3 rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; grep -E "passed|failed" ${O}_pytest.log | tail -n 2
python - <<'PY'
import json
for f in ['cfg2','cfg3','predict_bf16x3']:
    try:
        d=json.loads(open('gpurun_out/r02_c35_bench_%s.json'%f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], r['bound'], r['kernel'], round(r['frac'],4), {k: round(v,3) for k,v in r['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
