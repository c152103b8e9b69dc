#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
tail -3 gpurun_out/c12_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/c12_bench.json
