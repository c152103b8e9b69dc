#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c13_pytest.log 2>&1; tail -3 gpurun_out/c13_pytest.log
timeout 200 python scratch/msm_time.py both 0 2>&1 | tail -2
ZA_MSM_REDUCE_OLD=1 timeout 200 python scratch/msm_time.py both 0 2>&1 | tail -2
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -8
