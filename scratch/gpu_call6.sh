#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "create_proof" ) > gpurun_out/c6_pytest.log 2>&1; tail -3 gpurun_out/c6_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; python scratch/show_bench.py gpurun_out/c6_bench.json 2>/dev/null || head -c 600 gpurun_out/c6_bench.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/c6_acc_g1t -f python scratch/prof_target.py g1t > gpurun_out/c6_ncu_g1t.log 2>&1
