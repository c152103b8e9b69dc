#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 python scratch/dbg_prove.py 2>&1 | tail -1
ZA_H_OVERLAP=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -2
ZA_H_OVERLAP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "create_proof or synthetic_pk" 2>&1 | tail -2
