#!/bin/bash
# window index on the device: parity vs the oracle and the reference's golden vectors; CLI tests with the device index
mkdir -p gpurun_out
O=gpurun_out/r02_c29
( timeout 400 python -m pytest tests/test_gpu_window_index.py -m gpu -q -x --timeout 200 ) > ${O}_windex.log 2>&1; echo "windex rc=$?" > ${O}_rc.txt
( timeout 600 python -m pytest tests/test_gpu_cli.py -m gpu -q -x --timeout 300 ) > ${O}_cli.log 2>&1; echo "cli rc=$?" >> ${O}_rc.txt
cat ${O}_rc.txt; tail -n 25 ${O}_windex.log; tail -n 5 ${O}_cli.log
