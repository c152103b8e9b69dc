#!/bin/bash
# data-gathering call: GPU tests, proof timeline, ncu captures of the pair-round kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
tail -3 gpurun_out/c1_pytest.log
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py > gpurun_out/c1_timeline.log 2>&1
tail -30 gpurun_out/c1_timeline.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round -s 4 -c 2 -o gpurun_out/c1_pair_g2 -f python scratch/prof_target.py g2t > gpurun_out/c1_ncu_g2.log 2>&1
ZA_MSM_ROUNDS=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round -s 3 -c 2 -o gpurun_out/c1_pair_g1 -f python scratch/prof_target.py g1t > gpurun_out/c1_ncu_g1.log 2>&1
ls -la gpurun_out
