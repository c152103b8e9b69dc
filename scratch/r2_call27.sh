#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 400 python bench.py --steps 10 --warmup 3 --no-sub > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c27_bench.json; tail -3 gpurun_out/r2c27_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c27_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','profiled_ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_step']['frac'])
PY
