#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 120 python scratch/r2_proof_time.py default
  ZA_G2_INLINE=1 timeout 120 python scratch/r2_proof_time.py g2inline
  ZA_MSM_CHUNKS_PER_SLOT=4 timeout 120 python scratch/r2_proof_time.py chunks4
  ZA_MSM_CHUNKS_PER_SLOT=8 timeout 120 python scratch/r2_proof_time.py chunks8
  ZA_G2_INLINE=1 ZA_MSM_CHUNKS_PER_SLOT=4 timeout 120 python scratch/r2_proof_time.py g2inline_chunks4
  ZA_G2_INLINE=1 ZA_H_INLINE=1 timeout 120 python scratch/r2_proof_time.py allinline
  ZA_H_INLINE=1 timeout 120 python scratch/r2_proof_time.py hinline
  ZA_MSM_MERGE=0 ZA_G2_INLINE=1 timeout 120 python scratch/r2_proof_time.py nomerge_g2inline ) > gpurun_out/r2c18_proof.log 2>&1
grep "^\[" gpurun_out/r2c18_proof.log
