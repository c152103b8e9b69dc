#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( ZA_G2_SM=1 ZA_MSM_PAIR_LP=16 timeout 200 python scratch/r2_variant_time.py g2sm_lp16 2>&1 | grep "^\[" | grep "G2\|proof 2"
  ZA_G2_SM=1 ZA_MSM_ROUNDS=4 timeout 200 python scratch/r2_variant_time.py g2sm_r4 2>&1 | grep "^\[" | grep "G2\|proof 2"
  ZA_G2_SM=1 ZA_MSM_ROUNDS=5 timeout 200 python scratch/r2_variant_time.py g2sm_r5 2>&1 | grep "^\[" | grep "G2\|proof 2" ) > gpurun_out/r2c35_g2.log 2>&1
cat gpurun_out/r2c35_g2.log
