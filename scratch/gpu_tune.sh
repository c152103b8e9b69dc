#!/bin/bash
cd $GRAFT_REPO_ROOT
exec < /dev/null
for k in 1 2 3 4; do echo "CHUNKS_PER_SLOT=$k"; ZA_MSM_CHUNKS_PER_SLOT=$k timeout 100 python scratch/dbg_prove.py 2>&1 | tail -2 | head -1; done
echo "PAIR_LP=16"; ZA_MSM_PAIR_LP=16 timeout 100 python scratch/dbg_prove.py 2>&1 | tail -2 | head -1
