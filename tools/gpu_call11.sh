#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "allpairs" > gpurun_out/r2c11_pytest_ap.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_pytest_ap.log
tail -n 6 gpurun_out/r2c11_pytest_ap.log | cut -c1-300
if grep -q "rc=0" gpurun_out/r2c11_pytest_ap.log; then
  timeout 300 python tools/allpairs_bench.py > gpurun_out/r2c11_allpairs.log 2>&1; cat gpurun_out/r2c11_allpairs.log | cut -c1-400
  timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_allpairs_topk_equals_float64_ranking > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log
  tail -n 4 gpurun_out/r2c11_pytest.log | cut -c1-300
fi
