#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
timeout 400 python bench.py --steps 10 --warmup 3 --no-sub > gpurun_out/r2c37_bench_n1.json 2> gpurun_out/r2c37_bench_n1.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c37_bench_n1.json; tail -3 gpurun_out/r2c37_bench_n1.err
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c37_bench_n$N.json 2> gpurun_out/r2c37_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c37_bench_n$N.json; tail -3 gpurun_out/r2c37_bench_n$N.err
