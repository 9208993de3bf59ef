#!/bin/bash
mkdir -p gpurun_out
TFR_SHARDED_EXCHANGE=allreduce_graph TFR_SHARDED_SCALE=200 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c15_sharded_small_graph.json 2> gpurun_out/r2c15_sharded_small_graph.err; echo "rc=$?" >> gpurun_out/r2c15_sharded_small_graph.err
tail -n 3 gpurun_out/r2c15_sharded_small_graph.err | cut -c1-300
if grep -q '"pass": true' gpurun_out/r2c15_sharded_small_graph.json; then
TFR_SHARDED_EXCHANGE=allreduce_graph timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c15_sharded_full_graph.json 2> gpurun_out/r2c15_sharded_full_graph.err; echo "rc=$?" >> gpurun_out/r2c15_sharded_full_graph.err
tail -n 2 gpurun_out/r2c15_sharded_full_graph.err | cut -c1-300
fi
python - <<'PY'
import json
for f in ("small_graph","full_graph"):
    try:
        d=json.load(open("gpurun_out/r2c15_sharded_%s.json"%f))
        print(f, "ms/step %.3f"%d["ms_per_step"], "pass %.3f"%d["roofline"]["launch_ms"], "step-pass %.3f"%d["comm"]["step_minus_pass_ms"], "parity", (d.get("parity_check") or {}).get("pass"), "e2e", d["e2e"]["value"])
    except Exception as e: print(f, "no result", e)
PY
