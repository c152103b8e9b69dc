#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c3_smi.txt
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
tail -15 gpurun_out/r2c3_pytest.log
