#!/bin/bash
mkdir -p gpurun_out
for cfg in "TFR_TL_EVERY_CTA=1" "TFR_TL_EVERY_CTA=1 TFR_STREAM_DYNAMIC=1" "TFR_STREAM_DYNAMIC=0" "TFR_STREAM_DYNAMIC=1"; do
  echo "=== $cfg" >> gpurun_out/r2c9_timeline.log
  env $cfg timeout 300 python tools/timeline.py ml25m_d128_b65536 >> gpurun_out/r2c9_timeline.log 2>&1
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train_step or stream_graph or full_size" > gpurun_out/r2c9_pytest_static.log 2>&1
TFR_STREAM_DYNAMIC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train_step or stream_graph or full_size" > gpurun_out/r2c9_pytest_dynamic.log 2>&1
cat gpurun_out/r2c9_timeline.log; tail -2 gpurun_out/r2c9_pytest_static.log gpurun_out/r2c9_pytest_dynamic.log
bash tools/gpu_call8.sh
