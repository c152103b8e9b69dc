#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
N=2
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2c42_bench_n$N.json 2> gpurun_out/r2c42_bench_n$N.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c42_bench_n2.json').read().strip().splitlines()[-1])
print(2, round(d['value'],3), d['roofline']['kernel'][:40], round(d['roofline']['share_of_step'],3), round(d['roofline_g2']['share_of_step'],3))
PY
tail -2 gpurun_out/r2c42_bench_n2.err
