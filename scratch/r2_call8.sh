#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "pair_rounds or multiexp or create_proof or staged or table" > gpurun_out/r2c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c8_pytest.log
tail -5 gpurun_out/r2c8_pytest.log
timeout 600 python scratch/r2_variant_time.py lanepair > gpurun_out/r2c8_lp.log 2>&1; grep "^\[" gpurun_out/r2c8_lp.log
ZA_G2_LANEPAIR=0 timeout 600 python scratch/r2_variant_time.py single > gpurun_out/r2c8_single.log 2>&1; grep "^\[" gpurun_out/r2c8_single.log
