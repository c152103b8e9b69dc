#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c36_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c36_pytest.log
grep -v "^\[za" gpurun_out/r2c36_pytest.log | tail -4
SEL="multiexp_matches_oracle_uniform or multiexp_witness_like or multiexp_duplicate or create_proof_mul_chain or babyadd or pair_rounds_exceptional or staged_prove"
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2c36_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|at void" gpurun_out/r2c36_memcheck.log | head -6
timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multiexp_matches_oracle_uniform or pair_rounds_exceptional or create_proof_mul_chain" > gpurun_out/r2c36_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|at void" gpurun_out/r2c36_racecheck.log | head -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-sub > gpurun_out/r2c36_bench.json 2> gpurun_out/r2c36_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c36_bench.json | head -4
