#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "ntt or h_poly or create_proof or staged" > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -5 gpurun_out/r2c2_pytest.log
for k in 9 10 11; do ZA_NTT_MAXK=$k timeout 300 python scratch/r2_ntt_time.py >> gpurun_out/r2c2_ntt.log 2>&1; done
grep "^\[" gpurun_out/r2c2_ntt.log
timeout 600 python scratch/r2_variant_time.py main > gpurun_out/r2c2_main.log 2>&1; grep "^\[" gpurun_out/r2c2_main.log
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py > gpurun_out/r2c2_timeline.log 2>&1
grep "timeline" gpurun_out/r2c2_timeline.log | tail -12
