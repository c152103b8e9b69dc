#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
SEL="multiexp_matches_oracle_uniform or multiexp_witness_like or multiexp_duplicate or create_proof_mul_chain or babyadd or pair_rounds_exceptional or staged_prove"
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/san_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|at void" gpurun_out/san_memcheck.log | head -8
timeout 170 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp_matches_oracle_uniform or babyadd or create_proof_mul_chain" > gpurun_out/san_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|at void" gpurun_out/san_racecheck.log | head -8
