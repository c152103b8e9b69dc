#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_rounds or msm or multiexp" ) > gpurun_out/c2_pytest.log 2>&1; tail -5 gpurun_out/c2_pytest.log
timeout 300 python scratch/msm_time.py both 0,3,4 2>&1 | tee gpurun_out/c2_msm_time.log
timeout 300 python scratch/dbg_prove.py 2>&1 | tail -4 | tee gpurun_out/c2_prove.log
