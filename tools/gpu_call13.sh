#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_c13_pytest.log 2>&1; echo "all rc=$?" > gpurun_out/r02_c13_rc.txt
LFMQ_TRACE_GEN=1 LFMQ_GEN_SPLIT=0 timeout 120 python tools/run_once.py --workload cfg3 --steps 2 > /dev/null 2> gpurun_out/r02_c13_gtrace.txt
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c13_cfg3.json 2> gpurun_out/r02_c13_cfg3.err
cat gpurun_out/r02_c13_rc.txt; tail -n 6 gpurun_out/r02_c13_pytest.log
python - <<'PY'
import json
for f in ('r02_c13_cfg3',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -E "bwd l=1 t=(16|24)" gpurun_out/r02_c13_gtrace.txt | head -2
