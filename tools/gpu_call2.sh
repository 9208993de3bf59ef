#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ring or train_step or full_size" > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
timeout 600 python tools/pass_bench.py ml25m_d128_b65536 30 \
  PASS_RING=0 \
  PASS_RING=1 \
  RING_L2_HINT=0 RING_L2_HINT=1 RING_L2_HINT=3 \
  RING_L2_HINT=2,RING_STAGES=3 RING_STAGES=5,SMEM_CARVEOUT=100 RING_STAGES=6,RING_STAGE_KB=16 RING_STAGES=8,RING_STAGE_KB=12 RING_STAGES=3,RING_STAGE_KB=36 RING_STAGES=2,RING_STAGE_KB=48 \
  RING_STAGES=4,RING_STAGE_KB=24,RING_THREADS=160 RING_THREADS=224 RING_THREADS=320 \
  RING_CTAS_PER_SM=1,RING_THREADS=544,RING_STAGES=8,RING_STAGE_KB=24 RING_STAGES=4,RING_STAGE_KB=48 RING_STAGES=6,RING_STAGE_KB=32 RING_THREADS=800,RING_STAGES=8,RING_STAGE_KB=24\
  RING_CTAS_PER_SM=3,RING_THREADS=224,RING_STAGES=4,RING_STAGE_KB=16 RING_CTAS_PER_SM=4,RING_THREADS=160,RING_STAGES=4,RING_STAGE_KB=12 \
  > gpurun_out/r2c2_pass_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_pass_bench.log
for cfg in "TFR_PASS_RING=0" "TFR_PASS_RING=1" "TFR_PASS_RING=1 TFR_SMEM_CARVEOUT=100" "TFR_PASS_RING=1 TFR_SMEM_CARVEOUT=85" "TFR_PASS_RING=1 TFR_RING_STAGES=3" "TFR_PASS_RING=1 TFR_RING_CTAS_PER_SM=1 TFR_RING_THREADS=544 TFR_RING_STAGES=5"; do
  echo "=== $cfg" >> gpurun_out/r2c2_timeline.log
  env $cfg timeout 300 python tools/timeline.py ml25m_d128_b65536 >> gpurun_out/r2c2_timeline.log 2>&1
done
tail -3 gpurun_out/r2c2_pytest.log; cat gpurun_out/r2c2_pass_bench.log gpurun_out/r2c2_timeline.log
