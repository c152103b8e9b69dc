#!/bin/bash
# round-2 evidence on one B200: GPU tests, smoke, launch list of a bench run, ncu --set full of the dominant kernel,
# full bench line, reference arm, config-5 sweeps
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
T=${1:-r2ev}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -v "^\[za" gpurun_out/${T}_pytest.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub > gpurun_out/${T}_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_g1_sm -s 1 -c 1 -o gpurun_out/${T}_acc_g1 python scratch/prof_target.py g1t > gpurun_out/${T}_ncu_acc.log 2>&1; tail -2 gpurun_out/${T}_ncu_acc.log
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/${T}_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; head -c 600 gpurun_out/${T}_bench_ref.json; echo
timeout 400 python scratch/sweep_config5.py 26 > gpurun_out/${T}_sweep_config5.md 2> gpurun_out/${T}_sweep_config5.err; tail -16 gpurun_out/${T}_sweep_config5.md
