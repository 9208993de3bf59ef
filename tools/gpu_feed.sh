#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or driver or range or smoke or session or sort" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -n 3 gpurun_out/r2g_pytest.log | cut -c1-300
for wl in ml25m_d128_b65536 ml1m_d15_b10000; do timeout 300 python tools/e2e_breakdown.py $wl 2>&1 | tail -n 13; done
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "also e2e", d["also"]["ms_per_step"], d["also"]["e2e"]["ms_per_step"])
PY
