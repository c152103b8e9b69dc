#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_rounds or msm or multiexp" ) > gpurun_out/c4_pytest.log 2>&1; tail -3 gpurun_out/c4_pytest.log
timeout 300 python scratch/msm_time.py both 0,3,4 2>&1 | tee gpurun_out/c4_msm_time.log
ZA_MSM_ROUNDS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"msm_pair_round|msm_accumulate" --csv --log-file gpurun_out/c4_g1_launches.csv python scratch/prof_target.py g1t > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"msm_pair_round|msm_accumulate" --csv --log-file gpurun_out/c4_g2_launches.csv python scratch/prof_target.py g2t > /dev/null 2>&1
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -9 | tee gpurun_out/c4_prove.log
