#!/bin/bash
mkdir -p gpurun_out
( timeout 240 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "generic or cfg3" ) > gpurun_out/r02_c11_generic.log 2>&1; echo "generic rc=$?" > gpurun_out/r02_c11_rc.txt
( LFMQ_GEN_SPLIT=1 timeout 200 python -m pytest tests/test_gpu_generic.py -m gpu -q -x -k "gradients or smaller or train_steps" ) > gpurun_out/r02_c11_split.log 2>&1; echo "split rc=$?" >> gpurun_out/r02_c11_rc.txt
LFMQ_GEN_SPLIT=0 timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c11_cfg3_split0.json 2> gpurun_out/r02_c11_cfg3_split0.err
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c11_cfg3_split1.json 2> gpurun_out/r02_c11_cfg3_split1.err
cat gpurun_out/r02_c11_rc.txt; tail -n 2 gpurun_out/r02_c11_generic.log; tail -n 2 gpurun_out/r02_c11_split.log
python - <<'PY'
import json
for f in ('r02_c11_cfg3_split0','r02_c11_cfg3_split1'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
