#!/bin/bash
# general path backward step: 64-unit tiles (256 CTAs per step), one or two CTAs per SM
mkdir -p gpurun_out
O=gpurun_out/r02_c28
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_$n.json 2> ${O}_$n.err
}
run main LFMQ_X=0
run bn64 LFMQ_GEN_BWD_BN64=1
run bn64_2cta LFMQ_GEN_BWD_BN64=1 LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_altB.so
run bn64_2cta_nosplit LFMQ_GEN_BWD_BN64=1 LFMQ_GEN_SPLIT=0 LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_altB.so
( LFMQ_GEN_BWD_BN64=1 LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_altB.so timeout 200 python -m pytest tests/test_gpu_generic.py -m gpu -q -x ) > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 ${O}_tests.log
python - <<'PY'
import json
for n in ('main','bn64','bn64_2cta','bn64_2cta_nosplit'):
    try:
        d=json.loads(open('gpurun_out/r02_c28_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
