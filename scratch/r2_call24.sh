#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c24_pytest.log
grep -v "^\[za" gpurun_out/r2c24_pytest.log | tail -6
( timeout 200 python scratch/r2_variant_time.py w2d
  ZA_MSM_W2D=0 timeout 200 python scratch/r2_variant_time.py old
  ZA_SHARD_PLAN=1 timeout 120 python scratch/r2_shard_time.py w2d_dev2 8 2
  ZA_SHARD_PLAN=1 ZA_MSM_W2D=0 timeout 120 python scratch/r2_shard_time.py old_dev2 8 2 ) 2>&1 | grep "^\[" > gpurun_out/r2c24_w2d.log
cat gpurun_out/r2c24_w2d.log
