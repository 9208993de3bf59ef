#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
cp gpurun_out/parity_stats.json gpurun_out/r2f_parity_stats.json 2>/dev/null
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2f_bench_ref.err
tail -n 4 gpurun_out/r2f_pytest.log | cut -c1-300; tail -n 2 gpurun_out/r2f_smoke.log; tail -n 1 gpurun_out/r2f_bench.err gpurun_out/r2f_bench_ref.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}); print("roofline", d["roofline"]["frac"], d["roofline"]["launch_ms"]); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("cpu", d["cpu_baseline"]["value"])
print("also", d["also"]["ms_per_step"], d["also"]["e2e"]["ms_per_step"]); print("fm", d.get("also_fm",{}).get("ms_per_step"))
ap=d.get("also_allpairs",{}); print({k:(round(v["ms"],2), round(v["tflops"],1), round(v["frac_of_tf32_peak"],3)) for k,v in ap.get("consumers",{}).items()} if "consumers" in ap else ap)
r=json.load(open("gpurun_out/r2f_bench_ref.json")); print("ref", r["value"], r["steps"], r["warmup"], r["config"]==d["config"])
PY
