#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3_eddsa" 2>&1 | tail -4
