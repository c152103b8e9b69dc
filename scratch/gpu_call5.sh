#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round -s 4 -c 1 -o gpurun_out/c5_pair_g2 -f python scratch/prof_target.py g2t > gpurun_out/c5_ncu_g2.log 2>&1
ZA_MSM_ROUNDS=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round -s 3 -c 1 -o gpurun_out/c5_pair_g1 -f python scratch/prof_target.py g1t > gpurun_out/c5_ncu_g1.log 2>&1
