#!/bin/bash
# persistent backward steps: two half-batch chains on two streams (default) vs one launch over all row tiles
mkdir -p gpurun_out
O=gpurun_out/r02_c50
run() { n=$1; shift; env "$@" timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_$n.json 2> ${O}_$n.err; }
run split1 LFMQ_X=0
run split0 LFMQ_GEN_SPLIT=0
run split1b LFMQ_X=0
python - <<'PY'
import json
for n in ('split1','split0','split1b'):
    try:
        d=json.loads(open('gpurun_out/r02_c50_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
