#!/bin/bash
# same-box A/B: current build vs the c37 build (before the multicast option and the cached dropout bits), cfg3, interleaved
mkdir -p gpurun_out
O=gpurun_out/r02_c43
for i in 1 2; do
  timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cur$i.json 2> ${O}_cur$i.err
  LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_c37.so timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_c37_$i.json 2> ${O}_c37_$i.err
done
python - <<'PY'
import json
for n in ('cur1','c37_1','cur2','c37_2'):
    try:
        d=json.loads(open('gpurun_out/r02_c43_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
