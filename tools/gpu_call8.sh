#!/bin/bash
mkdir -p gpurun_out
LFMQ_DEBUG_SYNC=1 timeout 60 python -u tools/dbg_generic.py > gpurun_out/r02_c8_dbg.log 2>&1; echo "rc=$?" >> gpurun_out/r02_c8_dbg.log
tail -n 12 gpurun_out/r02_c8_dbg.log | cut -c1-200
