#!/bin/bash
# first layer's db from the dW GEMM (constant-one padding column) instead of a column-sum pass over dz
mkdir -p gpurun_out
O=gpurun_out/r02_c48
( timeout 300 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py tests/test_gpu_bf16.py -m gpu -q -x --timeout 150 ) > ${O}_tests.log 2>&1; echo "tests rc=$?" > ${O}_rc.txt
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cfg3.json 2> ${O}_cfg3.err; echo "cfg3 rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 12 ${O}_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c48_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
