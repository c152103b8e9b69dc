#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( for v in g2pf1 g2pf2; do
    ZA_G2_SM=1 ZA_B200_SO=$PWD/za_b200/variants/libza_b200_$v.so timeout 200 python scratch/r2_variant_time.py $v 2>&1 | grep "^\[" | grep "G2\|proof 2"
  done
  ZA_G2_SM=1 timeout 200 python scratch/r2_variant_time.py g2pf0 2>&1 | grep "^\[" | grep "G2\|proof 2"
  ZA_G2_SM=1 ZA_MSM_ROUNDS=4 timeout 200 python scratch/r2_variant_time.py g2pf0_r4 2>&1 | grep "^\[" | grep "G2\|proof 2" ) > gpurun_out/r2c33_g2pf.log 2>&1
cat gpurun_out/r2c33_g2pf.log
