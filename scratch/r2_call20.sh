#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 600 python -m pytest tests/test_gpu_parity.py -k "staged" -m gpu -x -q ) > gpurun_out/r2c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c20_pytest.log
grep -v "^\[za" gpurun_out/r2c20_pytest.log | tail -5
for f in 8 4 2; do
  ZA_MSM_RED_FANIN=$f timeout 200 python scratch/r2_variant_time.py fan$f 2>&1 | grep "^\["
  ZA_MSM_RED_FANIN=$f timeout 120 python scratch/r2_shard_time.py fan$f 8 1 2>&1 | grep "^\["
done > gpurun_out/r2c20_fanin.log 2>&1
cat gpurun_out/r2c20_fanin.log
