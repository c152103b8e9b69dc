#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 600 python -m pytest tests/test_prover_multi.py tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/r2c29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c29_pytest.log
grep -v "^\[za" gpurun_out/r2c29_pytest.log | tail -6
