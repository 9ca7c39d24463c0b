#!/bin/bash
# last seconds of the round's GPU budget: smoke + the handler/engine GPU tests on the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c9_smoke.log 2>&1; tail -1 gpurun_out/c9_smoke.log
timeout 70 python -m pytest tests/test_shm_handler.py tests/test_engines.py tests/test_gpu_r02.py -m gpu -q -x > gpurun_out/c9_pytest.log 2>&1; grep -E "passed|failed|FAILED|ERROR" gpurun_out/c9_pytest.log | tail -3
