#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14_g1_launches.csv python scratch/prof_target.py g1t > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14_g2_launches.csv python scratch/prof_target.py g2t > /dev/null 2>&1
