#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c4_bench_n1.json 2> gpurun_out/r2c4_bench_n1.err; echo "n1 rc=$?"
tail -5 gpurun_out/r2c4_bench_n1.err
python scratch/show_bench.py gpurun_out/r2c4_bench_n1.json 2>/dev/null | head -60
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c4_bench_n2.json 2> gpurun_out/r2c4_bench_n2.err; echo "n2 rc=$?"
tail -5 gpurun_out/r2c4_bench_n2.err
python scratch/show_bench.py gpurun_out/r2c4_bench_n2.json 2>/dev/null | head -40
