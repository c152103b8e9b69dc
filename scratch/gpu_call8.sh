#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c8_pytest.log 2>&1; tail -3 gpurun_out/c8_pytest.log
ZA_MSM_ACC_SM=0 timeout 200 python scratch/msm_time.py g1 0 2>&1 | tail -2
ZA_MSM_ACC_SM=1 timeout 200 python scratch/msm_time.py g1 0,3 2>&1 | tail -3
timeout 300 python scratch/dbg_prove.py 2>&1 | tail -3
