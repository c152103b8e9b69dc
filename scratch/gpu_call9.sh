#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp_duplicate or multiexp_errors or linearity_large or example_circuit" > gpurun_out/c9_sanitizer.log 2>&1
grep -v "^$" gpurun_out/c9_sanitizer.log | head -60
