#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c5_bench_n$N.json 2> gpurun_out/r2c5_bench_n$N.err; echo "n$N rc=$?"
tail -3 gpurun_out/r2c5_bench_n$N.err
python scratch/show_bench.py gpurun_out/r2c5_bench_n$N.json 2>/dev/null | head -40
