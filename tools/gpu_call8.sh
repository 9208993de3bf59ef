#!/bin/bash
# profiles of round 2: launch list of the bench command, full captures of the step's kernels and of the all-pairs kernel
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-also --cpu-steps 1 > gpurun_out/r2c8_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c8_launches.csv python bench.py --steps 4 --warmup 3 --no-also --cpu-steps 1 > gpurun_out/r2c8_ncu_launches.log 2>&1
python tools/prof_step.py ml25m_d128_b65536 4 > gpurun_out/r2c8_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'adam_stream_multi|segsum_tiles|segsum_fixup|dedup_sort' --launch-skip 8 -c 4 -o gpurun_out/r2c8_step python tools/prof_step.py ml25m_d128_b65536 4 > gpurun_out/r2c8_ncu_step.log 2>&1
python tools/prof_allpairs.py > gpurun_out/r2c8_ap_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'allpairs_tc_kernel' -c 3 -o gpurun_out/r2c8_allpairs python tools/prof_allpairs.py > gpurun_out/r2c8_ncu_ap.log 2>&1
ls -la gpurun_out/r2c8_*; tail -3 gpurun_out/r2c8_ncu_step.log gpurun_out/r2c8_ncu_ap.log
