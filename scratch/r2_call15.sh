#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log
grep -v "^\[za" gpurun_out/r2c15_pytest.log | tail -6
( timeout 120 python scratch/r2_shard_time.py newrule 8 1
  ZA_MSM_ROUNDS=1 timeout 120 python scratch/r2_shard_time.py rounds1 8 1
  ZA_MSM_ROUNDS=3 timeout 120 python scratch/r2_shard_time.py rounds3 8 1
  ZA_MSM_ROUNDS=2 ZA_MSM_PAIR_LP=16 timeout 120 python scratch/r2_shard_time.py rounds2lp16 8 1
  ZA_MSM_ROUNDS=2 timeout 120 python scratch/r2_shard_time.py rounds2 4 1
  ZA_MSM_ROUNDS=3 timeout 120 python scratch/r2_shard_time.py rounds3 4 1
  ZA_MSM_ROUNDS=3 timeout 120 python scratch/r2_shard_time.py rounds3 2 1
  timeout 120 python scratch/r2_shard_time.py default 8 0 ) > gpurun_out/r2c15_shard.log 2>&1
grep "^\[" gpurun_out/r2c15_shard.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-sub > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c15_bench.json
