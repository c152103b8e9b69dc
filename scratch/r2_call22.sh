#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c22_pytest.log
grep -v "^\[za" gpurun_out/r2c22_pytest.log | tail -6
