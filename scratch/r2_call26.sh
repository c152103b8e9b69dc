#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
SEL="multiexp_matches_oracle_uniform or multiexp_witness_like or multiexp_duplicate or create_proof_mul_chain or babyadd or pair_rounds_exceptional or staged_prove"
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2c26_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|at void" gpurun_out/r2c26_memcheck.log | head -8
ZA_NTT_SM=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or h_poly" > gpurun_out/r2c26_memcheck_nttsm.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|at void" gpurun_out/r2c26_memcheck_nttsm.log | head -8
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp_matches_oracle_uniform or babyadd or create_proof_mul_chain" > gpurun_out/r2c26_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|at void" gpurun_out/r2c26_racecheck.log | head -8
ZA_NTT_SM=1 timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "h_poly" > gpurun_out/r2c26_racecheck_nttsm.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|at void" gpurun_out/r2c26_racecheck_nttsm.log | head -8
timeout 500 python scratch/sweep_config5.py 26 > gpurun_out/r2c26_sweep_config5.md 2> gpurun_out/r2c26_sweep_config5.err; tail -18 gpurun_out/r2c26_sweep_config5.md; tail -2 gpurun_out/r2c26_sweep_config5.err
