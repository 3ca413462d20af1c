#!/bin/bash
# last verification of the committed build: smoke(), full GPU suite, the driver's bench line, cfg3 bench line
mkdir -p gpurun_out
O=gpurun_out/r02_c49
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke rc=$?" > ${O}_rc.txt
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_rc.txt
timeout 600 python bench.py > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err; echo "bench cfg2 rc=$?" >> ${O}_rc.txt
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_cfg3.json 2> ${O}_bench_cfg3.err; echo "bench cfg3 rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 3 ${O}_smoke.log; grep -E "passed|failed" ${O}_pytest.log | tail -n 2
python - <<'PY'
import json
for f in ['cfg2','cfg3']:
    try:
        d=json.loads(open('gpurun_out/r02_c49_bench_%s.json'%f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], r['bound'], r['kernel'], round(r['frac'],4), {k: round(v,3) for k,v in r['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
