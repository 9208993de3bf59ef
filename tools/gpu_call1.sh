#!/bin/bash
# round 2, GPU call 1: suite, ring-pass sweep, contention experiment
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 600 python tools/pass_bench.py ml25m_d128_b65536 30 \
  PASS_RING=0 \
  PASS_RING=1,RING_STAGES=5,RING_STAGE_KB=24,RING_THREADS=544 \
  RING_STAGES=4 RING_STAGES=3 \
  SMEM_CARVEOUT=100,RING_STAGES=6 RING_STAGES=8 \
  RING_STAGES=4,RING_STAGE_KB=48 RING_STAGES=8,RING_STAGE_KB=12 RING_STAGES=16,RING_STAGE_KB=12 RING_STAGES=12,RING_STAGE_KB=6 \
  RING_STAGES=8,RING_STAGE_KB=24,RING_THREADS=288 RING_THREADS=416 RING_THREADS=800 RING_THREADS=1024 \
  RING_THREADS=544,RING_L2_HINT=1 RING_L2_HINT=2 \
  RING_L2_HINT=0,RING_CTAS_PER_SM=2,RING_STAGES=4,RING_STAGE_KB=24,RING_THREADS=288 \
  RING_CTAS_PER_SM=2,RING_STAGES=8,RING_STAGE_KB=12,RING_THREADS=288 \
  SMEM_CARVEOUT=62,RING_CTAS_PER_SM=1,RING_STAGES=5,RING_STAGE_KB=24,RING_THREADS=544 \
  > gpurun_out/r2c1_pass_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c1_pass_bench.log
timeout 600 python tools/contention.py PASS_RING=0 PASS_RING=0,STREAM_CTAS_PER_SM=1,STREAM_THREADS=512 \
  PASS_RING=1,RING_STAGES=3,RING_STAGE_KB=24,RING_THREADS=544 \
  PASS_RING=1,SMEM_CARVEOUT=100,RING_STAGES=5 RING_STAGES=4,RING_THREADS=288 \
  > gpurun_out/r2c1_contention.log 2>&1; echo "rc=$?" >> gpurun_out/r2c1_contention.log
timeout 300 python tools/timeline.py ml25m_d128_b65536 > gpurun_out/r2c1_timeline.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 5 --no-also > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_pass_bench.log gpurun_out/r2c1_contention.log gpurun_out/r2c1_timeline.log
