#!/bin/bash
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log
tail -5 gpurun_out/r2c11_pytest.log
timeout 600 python scratch/r2_variant_time.py merged > gpurun_out/r2c11_merged.log 2>&1; grep "^\[" gpurun_out/r2c11_merged.log
ZA_MSM_MERGE=0 timeout 600 python scratch/r2_variant_time.py separate > gpurun_out/r2c11_sep.log 2>&1; grep "^\[" gpurun_out/r2c11_sep.log
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py > gpurun_out/r2c11_timeline.log 2>&1
grep "timeline" gpurun_out/r2c11_timeline.log | tail -6
