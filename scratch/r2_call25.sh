#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time ZA_NTT_SM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "ntt or h_poly or create_proof" -m gpu -x -q ) > gpurun_out/r2c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c25_pytest.log
grep -v "^\[za" gpurun_out/r2c25_pytest.log | tail -6
ZA_NTT_SM=1 timeout 200 python scratch/r2_ntt_time.py 2>&1 | sed "s/^/[sm] /" > gpurun_out/r2c25_ntt.log
timeout 200 python scratch/r2_ntt_time.py 2>&1 | sed "s/^/[reg] /" >> gpurun_out/r2c25_ntt.log
grep "2^20\|2^24\|2^16" gpurun_out/r2c25_ntt.log
