#!/bin/bash
# general-path backward step: A tile multicast over a cluster of the row tile's column CTAs (LFMQ_GEN_MCAST=0: off)
mkdir -p gpurun_out
O=gpurun_out/r02_c39
timeout 100 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline > ${O}_probe.json 2> ${O}_probe.err; rc=$?
echo "probe rc=$rc" > ${O}_rc.txt
if [ $rc -ne 0 ]; then cat ${O}_rc.txt; tail -n 5 ${O}_probe.err; exit 0; fi
( timeout 200 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x --timeout 100 ) > ${O}_tests.log 2>&1; echo "tests rc=$?" >> ${O}_rc.txt
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cfg3_mcast.json 2> ${O}_cfg3_mcast.err
LFMQ_GEN_MCAST=0 timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cfg3_nomcast.json 2> ${O}_cfg3_nomcast.err
LFMQ_GEN_PERSIST=0 timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > ${O}_cfg3_mcast_stepped.json 2> ${O}_cfg3_mcast_stepped.err
cat ${O}_rc.txt; tail -n 2 ${O}_tests.log
python - <<'PY'
import json
for n in ('mcast','nomcast','mcast_stepped'):
    try:
        d=json.loads(open('gpurun_out/r02_c39_cfg3_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
PY
