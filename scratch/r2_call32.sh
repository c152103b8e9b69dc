#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
ZA_G2_SM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round_g2sm -s 3 -c 3 -o gpurun_out/r2c32_g2sm python scratch/prof_target.py g2t > gpurun_out/r2c32_ncu.log 2>&1; tail -2 gpurun_out/r2c32_ncu.log
