#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time ZA_G2_SM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "multiexp or pair_rounds or create_proof or staged" -m gpu -x -q ) > gpurun_out/r2c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c31_pytest.log
grep -v "^\[za" gpurun_out/r2c31_pytest.log | tail -6
( ZA_G2_SM=1 timeout 200 python scratch/r2_variant_time.py g2sm
  timeout 200 python scratch/r2_variant_time.py default ) 2>&1 | grep "^\[" > gpurun_out/r2c31_g2sm.log
cat gpurun_out/r2c31_g2sm.log
