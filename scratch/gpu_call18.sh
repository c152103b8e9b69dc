#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "create_proof or staged or helper or synthetic_pk" ) > gpurun_out/c18_pytest.log 2>&1; tail -2 gpurun_out/c18_pytest.log
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -8
ZA_G2_INLINE=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -1
