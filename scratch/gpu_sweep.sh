#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 230 python scratch/sweep_config5.py 26 > gpurun_out/sweep_config5.md 2> gpurun_out/sweep_config5.err
cat gpurun_out/sweep_config5.md
