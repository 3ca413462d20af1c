#!/bin/bash
# final evidence pass: launch list of the cfg2 step, ncu --set full of the cluster kernels and of the persistent general-path kernels
mkdir -p gpurun_out
O=gpurun_out/r02_c36
LFMQ_BWD_PREFETCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file ${O}_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-loss-check > ${O}_ncu_bench.log 2>&1; echo "ncu list rc=$?" > ${O}_rc.txt
LFMQ_BWD_PREFETCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"lstm_bwd_tc_kernel|lstm_fwd_tc_kernel|wgrad_tc_kernel|head_tc_kernel" -c 12 -o ${O}_cfg2_full -f python tools/run_once.py --workload cfg2 --steps 3 > ${O}_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?" >> ${O}_rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_gemm_kernel -c 24 -o ${O}_cfg3_full -f python tools/run_once.py --workload cfg3 --steps 1 > ${O}_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload predict > ${O}_bench_predict_bf16.json 2> ${O}_bench_predict_bf16.err; echo "predict rc=$?" >> ${O}_rc.txt
timeout 300 python bench.py --workload batcher --steps 50 > ${O}_bench_batcher.json 2> ${O}_bench_batcher.err; echo "batcher rc=$?" >> ${O}_rc.txt
LFMQ_TRACE_BWD=1 timeout 120 python tools/time_steps.py --steps 2 --predict-batch 4096 > /dev/null 2> ${O}_btrace.txt
cat ${O}_rc.txt; ls -la gpurun_out/r02_c36*.ncu-rep; tail -n 1 ${O}_bench_predict_bf16.json | cut -c1-200
