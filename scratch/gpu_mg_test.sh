#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 100 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
