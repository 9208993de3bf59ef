#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c4_pytest.log
for cfg in "TFR_PASS_RING=0 TFR_PDL=0" "TFR_PASS_RING=0 TFR_PDL=1" "TFR_PASS_RING=0 TFR_PDL=1 TFR_SMEM_CARVEOUT=62" \
           "TFR_PASS_RING=1 TFR_PDL=1" "TFR_PASS_RING=1 TFR_PDL=1 TFR_RING_L2_HINT=3" "TFR_PASS_RING=1 TFR_PDL=1 TFR_RING_SLOT_MODE=3" \
           "TFR_PASS_RING=1 TFR_PDL=1 TFR_RING_SLOT_MODE=2 TFR_RING_L2_HINT=3" \
           "TFR_PASS_RING=1 TFR_PDL=1 TFR_RING_L2_HINT=3 TFR_RING_CTAS_PER_SM=4 TFR_RING_THREADS=192 TFR_RING_STAGES=4 TFR_RING_STAGE_KB=12"; do
  echo "=== $cfg" >> gpurun_out/r2c4_timeline.log
  env $cfg timeout 300 python tools/timeline.py ml25m_d128_b65536 >> gpurun_out/r2c4_timeline.log 2>&1
done
echo "=== ml1m PDL=1" >> gpurun_out/r2c4_timeline.log
timeout 300 python tools/timeline.py ml1m_d15_b10000 >> gpurun_out/r2c4_timeline.log 2>&1
echo "=== ml1m PDL=0" >> gpurun_out/r2c4_timeline.log
TFR_PDL=0 timeout 300 python tools/timeline.py ml1m_d15_b10000 >> gpurun_out/r2c4_timeline.log 2>&1
TFR_PASS_RING=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-also --cpu-steps 2 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
tail -3 gpurun_out/r2c4_pytest.log; cat gpurun_out/r2c4_timeline.log; cat gpurun_out/r2c4_bench.json
