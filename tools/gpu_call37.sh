#!/bin/bash
# EPI_STORE GEMMs with the column tiles of a row tile dispatched together (A tile shared in L2)
mkdir -p gpurun_out
O=gpurun_out/r02_c37
( timeout 200 python -m pytest tests/test_gpu_generic.py tests/test_gpu_baseline_shapes.py -m gpu -q -x --timeout 100 ) > ${O}_tests.log 2>&1; echo "tests rc=$?" > ${O}_rc.txt
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_cfg3.json 2> ${O}_cfg3.err; echo "cfg3 rc=$?" >> ${O}_rc.txt
timeout 300 ncu --set full --clock-control none -k regex:"tile_gemm_kernel<128, 2|gwgrad_kernel" -c 6 -o ${O}_store_wgrad -f python tools/run_once.py --workload cfg3 --steps 1 > ${O}_ncu.log 2>&1; echo "ncu rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 2 ${O}_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c37_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
