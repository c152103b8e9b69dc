#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 55 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --log-msm 20 --log-ntt 20 --log-setup 0 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c3_bench.json')); print(d['value'], d['submetrics'].get('config3_eddsa_mimc'))"
tail -2 gpurun_out/c3_bench.err
