#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "allpairs" > gpurun_out/r2c12_pytest_ap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c12_pytest_ap.log
tail -n 3 gpurun_out/r2c12_pytest_ap.log | cut -c1-300
timeout 300 python tools/allpairs_bench.py > gpurun_out/r2c12_allpairs.log 2>&1; cat gpurun_out/r2c12_allpairs.log | cut -c1-300
TFR_SHARDED_EXCHANGE=allreduce_graph TFR_SHARDED_SCALE=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c12_sharded_small_graph.json 2> gpurun_out/r2c12_sharded_small_graph.err; echo "rc=$?" >> gpurun_out/r2c12_sharded_small_graph.err
tail -n 3 gpurun_out/r2c12_sharded_small_graph.err | cut -c1-300; cut -c1-200 gpurun_out/r2c12_sharded_small_graph.json
if grep -q '"pass": true' gpurun_out/r2c12_sharded_small_graph.json; then
TFR_SHARDED_EXCHANGE=allreduce_graph timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c12_sharded_full_graph.json 2> gpurun_out/r2c12_sharded_full_graph.err; echo "rc=$?" >> gpurun_out/r2c12_sharded_full_graph.err
tail -n 2 gpurun_out/r2c12_sharded_full_graph.err | cut -c1-300
fi
python - <<'PY'
import json
for f in ("small_graph","full_graph"):
    try:
        d=json.load(open("gpurun_out/r2c12_sharded_%s.json"%f))
        print(f, "ms/step %.3f"%d["ms_per_step"], "pass %.3f"%d["roofline"]["launch_ms"], "step-pass %.3f"%d["comm"]["step_minus_pass_ms"], "parity", (d.get("parity_check") or {}).get("pass"))
    except Exception as e: print(f, "no result", e)
PY
