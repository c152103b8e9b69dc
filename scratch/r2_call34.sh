#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time ZA_G2_SM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "multiexp or pair_rounds or create_proof_mul" -m gpu -x -q ) > gpurun_out/r2c34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c34_pytest.log
grep -v "^\[za" gpurun_out/r2c34_pytest.log | tail -4
ZA_G2_SM=1 timeout 200 python scratch/r2_variant_time.py g2sm_pipe 2>&1 | grep "^\[" | grep "G2\|proof 2" > gpurun_out/r2c34_g2.log
cat gpurun_out/r2c34_g2.log
