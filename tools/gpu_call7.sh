#!/bin/bash
# 2 GPUs: single-GPU suite on GPU 0 first, then the sharded bench at small scale (parity check vs the oracle over real NCCL)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c7_pytest.log
cp gpurun_out/parity_stats.json gpurun_out/r2c7_parity_stats.json 2>/dev/null
for ex in a2a allreduce; do
TFR_SHARDED_EXCHANGE=$ex TFR_SHARDED_SCALE=200 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c7_sharded_small_$ex.json 2> gpurun_out/r2c7_sharded_small_$ex.err; echo "rc=$?" >> gpurun_out/r2c7_sharded_small_$ex.err
done
for ex in a2a allreduce; do
TFR_SHARDED_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c7_sharded_full_$ex.json 2> gpurun_out/r2c7_sharded_full_$ex.err; echo "rc=$?" >> gpurun_out/r2c7_sharded_full_$ex.err
done
tail -25 gpurun_out/r2c7_pytest.log | cut -c1-300
for f in gpurun_out/r2c7_sharded_*.json; do echo "== $f"; cut -c1-3000 $f; tail -5 ${f%.json}.err | cut -c1-400; done
