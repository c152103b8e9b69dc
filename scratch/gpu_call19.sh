#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c19_pytest.log 2>&1; tail -2 gpurun_out/c19_pytest.log
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; tail -2 gpurun_out/c19_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/c19_bench.json
