#!/bin/bash
# round-end evidence: all GPU tests, smoke, launch list of a bench run, full bench line (cpu baseline included)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c17_pytest.log 2>&1; tail -4 gpurun_out/c17_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c17_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/c17_bench_under_ncu.log 2>&1
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/c17_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/c17_bench_ref.json 2> gpurun_out/c17_bench_ref.err; head -c 700 gpurun_out/c17_bench_ref.json
