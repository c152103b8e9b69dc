#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c10_pytest.log 2>&1; tail -2 gpurun_out/c10_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 1 -c 1 -o gpurun_out/c10_acc_sm -f python scratch/prof_target.py g1t > gpurun_out/c10_ncu.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp or create_proof or pair_rounds_exceptional or staged or helper_prove or generate_parameters_example" > gpurun_out/c10_sanitizer.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|at void" gpurun_out/c10_sanitizer.log | head -10
