#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -u tools/dbg_generic.py > gpurun_out/r02_c9_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c9_dbg.log
( timeout 240 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "generic or cfg3 or bf16x3 or tensor_core" ) > gpurun_out/r02_c9_generic.log 2>&1; echo "generic rc=$?" >> gpurun_out/r02_c9_rc.txt
( LFMQ_GEN_DUAL=1 timeout 200 python -m pytest tests/test_gpu_generic.py -m gpu -q -x ) > gpurun_out/r02_c9_generic_dual.log 2>&1; echo "dual rc=$?" >> gpurun_out/r02_c9_rc.txt
LFMQ_TRACE_GEN=1 timeout 120 python tools/run_once.py --workload cfg3 --steps 2 > /dev/null 2> gpurun_out/r02_c9_gtrace.txt
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c9_cfg3.json 2> gpurun_out/r02_c9_cfg3.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c9_cfg2.json 2> gpurun_out/r02_c9_cfg2.err
tail -n 2 gpurun_out/r02_c9_dbg.log; cat gpurun_out/r02_c9_rc.txt; tail -n 2 gpurun_out/r02_c9_generic.log; tail -n 2 gpurun_out/r02_c9_generic_dual.log
python - <<'PY'
import json
for f in ('r02_c9_cfg3','r02_c9_cfg2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()}, d.get('loss_check',{}).get('max_rel_diff'))
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -E "fwd l=1 t=(16)|bwd l=1 t=(16)" gpurun_out/r02_c9_gtrace.txt | head -4
