#!/bin/bash
# round 2, GPU call 5: general path with PDL + 256x256 weight-gradient tiles; per-step clock64 trace
mkdir -p gpurun_out
( LFMQ_GEN_DUAL=1 timeout 600 python -m pytest tests/test_gpu_generic.py -m gpu -q -x ) > gpurun_out/r02_c5_generic_dual.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "generic or cfg3" ) > gpurun_out/r02_c5_generic.log 2>&1
LFMQ_TRACE_GEN=1 timeout 300 python tools/run_once.py --workload cfg3 --steps 2 > /dev/null 2> gpurun_out/r02_c5_gtrace.txt
LFMQ_GEN_PDL=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_cfg3_pdl0.json 2> gpurun_out/r02_c5_cfg3_pdl0.err
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_cfg3_pdl1.json 2> gpurun_out/r02_c5_cfg3_pdl1.err
LFMQ_GEN_DUAL=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_cfg3_dual0.json 2> gpurun_out/r02_c5_cfg3_dual0.err
tail -n 3 gpurun_out/r02_c5_generic_dual.log; tail -n 3 gpurun_out/r02_c5_generic.log
python - <<'PY'
import json
for f in ('r02_c5_cfg3_pdl0','r02_c5_cfg3_pdl1','r02_c5_cfg3_dual0'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -E "fwd l=1 t=(8|16|24)|bwd l=1 t=(8|16|24)|fwd l=0 t=16|bwd l=0 t=16" gpurun_out/r02_c5_gtrace.txt | head -12
