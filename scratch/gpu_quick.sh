#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 35 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --log-msm 20 --log-ntt 20 --log-setup 14 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c4_bench.json')); print(d['value'], d['roofline']['frac'], list(d['submetrics'].keys()), d['submetrics']['real_key_pipeline'].get('proof_verifies'), d['submetrics']['config3_eddsa_mimc'].get('proof_verifies'))"
tail -1 gpurun_out/c4_bench.err
