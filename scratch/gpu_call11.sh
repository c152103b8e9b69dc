#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp or pair_rounds" ) > gpurun_out/c11_pytest.log 2>&1; tail -2 gpurun_out/c11_pytest.log
timeout 200 python scratch/msm_time.py g2 0,4 2>&1 | tail -3
timeout 300 python scratch/dbg_prove.py 2>&1 | tail -2
