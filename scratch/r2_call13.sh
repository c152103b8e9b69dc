#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 python -m pytest tests/test_za2c_gpu.py -x -q -m gpu > gpurun_out/r2c13_za2c.log 2>&1; echo "za2c rc=$?" >> gpurun_out/r2c13_za2c.log
grep -v "^\[za" gpurun_out/r2c13_za2c.log | tail -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 6 -c 6 -o gpurun_out/r2c13_ntt python scratch/prof_target.py hpoly > gpurun_out/r2c13_ncu.log 2>&1; tail -2 gpurun_out/r2c13_ncu.log
timeout 200 python scratch/r2_ntt_time.py > gpurun_out/r2c13_ntt_time.log 2>&1; grep "2^20\|2^24" gpurun_out/r2c13_ntt_time.log
