#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 200 python scratch/sweep_config5_multi.py 20 1 16 > gpurun_out/r2c17_sweep_test.md 2> gpurun_out/r2c17_sweep_test.err; tail -5 gpurun_out/r2c17_sweep_test.md; tail -3 gpurun_out/r2c17_sweep_test.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-sub > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c17_bench.json
timeout 200 python scratch/sweep_config5_multi.py 26 1 26 > gpurun_out/r2c17_sweep_26.md 2> gpurun_out/r2c17_sweep_26.err; tail -2 gpurun_out/r2c17_sweep_26.md; tail -3 gpurun_out/r2c17_sweep_26.err
