#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c38_bench_n$N.json 2> gpurun_out/r2c38_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c38_bench_n$N.json
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c38_bench_n$N.json 2> gpurun_out/r2c38_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c38_bench_n$N.json
ZA_DEBUG_TIMELINE=1 timeout 120 python scratch/r2_prover_tl.py 8 2> gpurun_out/r2c38_tl8.log | tail -1
grep "proof [0-9]*:" gpurun_out/r2c38_tl8.log | tail -3
grep "collected" gpurun_out/r2c38_tl8.log | tail -8
