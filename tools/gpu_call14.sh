#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "allpairs" > gpurun_out/r2c14_pytest_ap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c14_pytest_ap.log
tail -n 3 gpurun_out/r2c14_pytest_ap.log | cut -c1-300
timeout 300 python tools/ap_stage_timing.py 2>&1 | tail -4
timeout 300 python tools/allpairs_bench.py 2>&1 | cut -c1-200
