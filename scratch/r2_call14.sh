#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 python -m pytest tests/test_za2c_gpu.py -x -q -m gpu > gpurun_out/r2c14_za2c.log 2>&1; echo "za2c rc=$?" >> gpurun_out/r2c14_za2c.log
grep -v "^\[za" gpurun_out/r2c14_za2c.log | tail -4
for v in nttmb3 nttunr nttmb3unr; do
  ZA_B200_SO=$PWD/za_b200/variants/libza_b200_$v.so timeout 200 python scratch/r2_ntt_time.py 2>&1 | grep "2^20\|fft 2^24" | sed "s/^/[$v] /"
done > gpurun_out/r2c14_ntt_variants.log 2>&1
cat gpurun_out/r2c14_ntt_variants.log
( ZA_DEBUG_TIMELINE=1 timeout 120 python scratch/r2_shard_time.py default 8 1
  timeout 120 python scratch/r2_shard_time.py default 8 1
  ZA_G2_INLINE=1 timeout 120 python scratch/r2_shard_time.py g2inline 8 1
  ZA_MSM_ROUNDS=0 timeout 120 python scratch/r2_shard_time.py rounds0 8 1
  ZA_MSM_ROUNDS=2 timeout 120 python scratch/r2_shard_time.py rounds2 8 1
  ZA_MSM_MERGE=0 timeout 120 python scratch/r2_shard_time.py nomerge 8 1
  ZA_MSM_TABLE=14 timeout 120 python scratch/r2_shard_time.py table14 8 1
  ZA_MSM_TABLE=18 timeout 120 python scratch/r2_shard_time.py table18 8 1
  timeout 120 python scratch/r2_shard_time.py default 4 1
  timeout 120 python scratch/r2_shard_time.py default 2 1 ) > gpurun_out/r2c14_shard.log 2>&1
grep "^\[\|timeline" gpurun_out/r2c14_shard.log | grep -v "timeline\] [LA] " | tail -40
