#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
timeout 100 python scratch/msm_time.py both 0 2>&1 | tail -2
timeout 100 python scratch/dbg_prove.py 2>&1 | tail -2 | head -1
