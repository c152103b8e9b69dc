#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
N=8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c41_bench_n$N.json 2> gpurun_out/r2c41_bench_n$N.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c41_bench_n8.json').read().strip().splitlines()[-1])
print(8, round(d['value'],3), round(d['e2e']['value'],3), d['cpu_baseline'].get('proof_matches_gpu'), d['proof_sha256'][:12])
PY
