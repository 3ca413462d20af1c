#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_generic.py -m gpu -q -x ) > gpurun_out/r02_c16_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02_c16_rc.txt
( LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 200 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x ) > gpurun_out/r02_c16_tests_alt.log 2>&1; echo "alt tests rc=$?" >> gpurun_out/r02_c16_rc.txt
echo "== main (LATE_C1, ct reuse)" > gpurun_out/r02_c16_time.txt
for i in 1 2; do timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c16_time.txt 2>&1; done
echo "== alt (LATE_C0 + LATE_C1)" >> gpurun_out/r02_c16_time.txt
for i in 1 2; do LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so timeout 120 python tools/time_steps.py --steps 20 --predict-batch 4096 >> gpurun_out/r02_c16_time.txt 2>&1; done
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c16_cfg3.json 2> gpurun_out/r02_c16_cfg3.err
cat gpurun_out/r02_c16_rc.txt; tail -n 2 gpurun_out/r02_c16_tests.log; tail -n 2 gpurun_out/r02_c16_tests_alt.log; grep -E "==|train" gpurun_out/r02_c16_time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c16_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
PY
