#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/c16_bench_n$N.json 2> gpurun_out/c16_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/c16_bench_n$N.json
