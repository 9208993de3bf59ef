#!/bin/bash
mkdir -p gpurun_out
python tools/prof_allpairs.py > gpurun_out/r2c13_ap_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'allpairs_tc_kernel' --launch-skip 1 -c 2 -o gpurun_out/r2c13_allpairs python tools/prof_allpairs.py > gpurun_out/r2c13_ncu_ap.log 2>&1
ls -la gpurun_out/r2c13_*; tail -n 3 gpurun_out/r2c13_ncu_ap.log
