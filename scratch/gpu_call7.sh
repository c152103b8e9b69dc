#!/bin/bash
# round-end style evidence run: ncu of the dominant kernel (traffic), launch list of a bench run, full bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/c7_acc_g1t -f python scratch/prof_target.py g1t > gpurun_out/c7_ncu_g1t.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c7_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/c7_bench_under_ncu.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/c7_bench.json
