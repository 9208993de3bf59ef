#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
timeout 600 python tools/pass_bench.py ml25m_d128_b65536 30 \
  PASS_RING=0,SMEM_CARVEOUT=62 \
  PASS_RING=1,SMEM_CARVEOUT=100 \
  RING_SLOT_MODE=0 RING_SLOT_MODE=2 RING_SLOT_MODE=3 RING_SLOT_MODE=3,STREAM_COPY_ONLY=1 RING_SLOT_MODE=1,STREAM_COPY_ONLY=0 \
  RING_L2_HINT=0 RING_L2_HINT=1 RING_L2_HINT=3 RING_L2_HINT=2 \
  RING_STAGES=3 RING_STAGES=6,RING_STAGE_KB=16 RING_STAGES=8,RING_STAGE_KB=12 RING_STAGES=3,RING_STAGE_KB=36 RING_STAGES=2,RING_STAGE_KB=48 \
  RING_STAGES=4,RING_STAGE_KB=24,RING_THREADS=192 RING_THREADS=256 RING_THREADS=384 \
  RING_CTAS_PER_SM=1,RING_THREADS=576,RING_STAGES=8,RING_STAGE_KB=24 RING_STAGES=4,RING_STAGE_KB=48 RING_STAGES=6,RING_STAGE_KB=32 RING_THREADS=832,RING_STAGES=8,RING_STAGE_KB=24 \
  RING_CTAS_PER_SM=3,RING_THREADS=256,RING_STAGES=4,RING_STAGE_KB=16 RING_CTAS_PER_SM=4,RING_THREADS=192,RING_STAGES=4,RING_STAGE_KB=12 \
  > gpurun_out/r2c3_pass_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c3_pass_bench.log
for cfg in "TFR_PASS_RING=0 TFR_SMEM_CARVEOUT=62" "TFR_PASS_RING=1" "TFR_PASS_RING=1 TFR_RING_STAGES=3 TFR_SMEM_CARVEOUT=75" "TFR_PASS_RING=1 TFR_RING_CTAS_PER_SM=1 TFR_RING_THREADS=576 TFR_RING_STAGES=4 TFR_RING_STAGE_KB=48"; do
  echo "=== $cfg" >> gpurun_out/r2c3_timeline.log
  env $cfg timeout 300 python tools/timeline.py ml25m_d128_b65536 >> gpurun_out/r2c3_timeline.log 2>&1
done
echo "=== ml1m" >> gpurun_out/r2c3_timeline.log
timeout 300 python tools/timeline.py ml1m_d15_b10000 >> gpurun_out/r2c3_timeline.log 2>&1
tail -3 gpurun_out/r2c3_pytest.log; cat gpurun_out/r2c3_pass_bench.log gpurun_out/r2c3_timeline.log
