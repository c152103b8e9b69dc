#!/bin/bash
mkdir -p gpurun_out
# ncu: first lane-pair round (FIRST=true LP=32) and a dense round, --set full with source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msm_pair_round_g2lp_kernel -c 2 -o gpurun_out/r2c9_g2lp -f python scratch/prof_target.py g2t > gpurun_out/r2c9_ncu.log 2>&1
ncu -i gpurun_out/r2c9_g2lp.ncu-rep --page raw --csv > gpurun_out/r2c9_g2lp.raw.csv 2>/dev/null
ncu -i gpurun_out/r2c9_g2lp.ncu-rep --page source --csv > gpurun_out/r2c9_g2lp.src.csv 2>/dev/null
python scratch/ncu_brief.py gpurun_out/r2c9_g2lp.raw.csv
python scratch/src_hot.py gpurun_out/r2c9_g2lp.src.csv 0 40
