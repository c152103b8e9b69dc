#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
for pad in 0 122880 90000; do echo "pad $pad"; ZA_G2_SMEM_PAD=$pad ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py 2>&1 | tail -7 | grep -v "after H"; done
