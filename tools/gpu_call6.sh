#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c6_pytest.log
cp gpurun_out/parity_stats.json gpurun_out/r2c6_parity_stats.json 2>/dev/null
timeout 300 python tools/e2e_breakdown.py ml25m_d128_b65536 > gpurun_out/r2c6_e2e.log 2>&1
timeout 300 python tools/e2e_breakdown.py ml1m_d15_b10000 >> gpurun_out/r2c6_e2e.log 2>&1
timeout 600 python tools/allpairs_bench.py > gpurun_out/r2c6_allpairs.log 2>&1
timeout 900 python bench.py --steps 200 --warmup 5 --cpu-steps 2 > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err
tail -30 gpurun_out/r2c6_pytest.log | cut -c1-400; cat gpurun_out/r2c6_e2e.log gpurun_out/r2c6_allpairs.log; cat gpurun_out/r2c6_bench.json | cut -c1-6000
