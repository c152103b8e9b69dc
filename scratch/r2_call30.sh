#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round -s 3 -c 3 -o gpurun_out/r2c30_g2_rounds python scratch/prof_target.py g2t > gpurun_out/r2c30_ncu.log 2>&1; tail -2 gpurun_out/r2c30_ncu.log
timeout 300 ncu --set full --clock-control none -k regex:msm_accumulate_kernel -s 1 -c 1 -o gpurun_out/r2c30_g2_acc python scratch/prof_target.py g2t > gpurun_out/r2c30_ncu2.log 2>&1; tail -2 gpurun_out/r2c30_ncu2.log
