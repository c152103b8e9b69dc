#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( for d in 2 4 5 7 0; do
    ZA_SHARD_PLAN=1 timeout 120 python scratch/r2_shard_time.py plan_dev$d 8 $d
    ZA_SHARD_PLAN=1 ZA_H_AFTER_G2=1 timeout 120 python scratch/r2_shard_time.py plan_dev${d}_hafter 8 $d
  done
  ZA_SHARD_PLAN=1 timeout 120 python scratch/r2_shard_time.py plan_dev1 4 1
  ZA_SHARD_PLAN=1 ZA_H_AFTER_G2=1 timeout 120 python scratch/r2_shard_time.py plan_dev1_hafter 4 1
  ZA_SHARD_PLAN=1 timeout 120 python scratch/r2_shard_time.py plan_dev2 4 2
  ZA_SHARD_PLAN=1 ZA_H_AFTER_G2=1 timeout 120 python scratch/r2_shard_time.py plan_dev2_hafter 4 2
  ZA_H_AFTER_G2=1 timeout 120 python scratch/r2_proof_time.py hafter_n1 ) > gpurun_out/r2c21_shard.log 2>&1
grep "^\[\|^plan" gpurun_out/r2c21_shard.log
