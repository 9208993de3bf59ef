#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "allpairs or binary_metrics or ktm" > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c10_pytest.log
timeout 600 python tools/allpairs_bench.py > gpurun_out/r2c10_allpairs.log 2>&1
i=0
for cfg in "TFR_PASS_RING=0" "TFR_PASS_RING=1 TFR_RING_L2_HINT=3" "TFR_PASS_RING=1 TFR_RING_L2_HINT=3 TFR_RING_CTAS_PER_SM=1 TFR_RING_THREADS=576 TFR_RING_STAGES=4 TFR_RING_STAGE_KB=48"; do
  i=$((i+1))
  env $cfg TFR_SHARDED_EXCHANGE=allreduce timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2c10_sharded_$i.json 2> gpurun_out/r2c10_sharded_$i.err; echo "$cfg rc=$?" >> gpurun_out/r2c10_sharded_$i.err
done
tail -n 4 gpurun_out/r2c10_pytest.log; cat gpurun_out/r2c10_allpairs.log
for i in 1 2 3; do tail -n 1 gpurun_out/r2c10_sharded_$i.err; cut -c1-1800 gpurun_out/r2c10_sharded_$i.json; done
