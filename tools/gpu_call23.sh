#!/bin/bash
# evidence pass 1 of the final build: full GPU suite, bench lines of every workload, launch list, ncu --set full of the two recurrence kernels
mkdir -p gpurun_out
O=gpurun_out/r02_c23
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 ) > ${O}_pytest.log 2>&1; echo "pytest rc=$?" > ${O}_rc.txt
timeout 600 python bench.py > ${O}_bench_cfg2.json 2> ${O}_bench_cfg2.err; echo "bench cfg2 rc=$?" >> ${O}_rc.txt
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > ${O}_bench_cfg3.json 2> ${O}_bench_cfg3.err; echo "bench cfg3 rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload predict > ${O}_bench_predict_bf16.json 2> ${O}_bench_predict_bf16.err; echo "predict rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload predict --precision bf16x3 > ${O}_bench_predict_bf16x3.json 2> ${O}_bench_predict_bf16x3.err; echo "predict x3 rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload batcher --steps 50 > ${O}_bench_batcher.json 2> ${O}_bench_batcher.err; echo "batcher rc=$?" >> ${O}_rc.txt
LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> ${O}_btrace.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file ${O}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-loss-check > ${O}_ncu_bench.log 2>&1; echo "ncu list rc=$?" >> ${O}_rc.txt
LFMQ_BWD_PREFETCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_tc_kernel -c 3 -o ${O}_bwd_full -f python tools/run_once.py --workload cfg2 --steps 3 > ${O}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?" >> ${O}_rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"lstm_fwd_tc_kernel|wgrad_tc_kernel|head_tc_kernel" -c 9 -o ${O}_fwd_full -f python tools/run_once.py --workload cfg2 --steps 3 > ${O}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 4 ${O}_pytest.log
for f in cfg2 cfg3 predict_bf16 predict_bf16x3 batcher; do tail -n 1 ${O}_bench_$f.json | cut -c1-400; done
ls -la gpurun_out/*.ncu-rep | tail -3
