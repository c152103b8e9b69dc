#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/c15_pytest.log 2>&1; tail -2 gpurun_out/c15_pytest.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c15_bench_n2.json 2> gpurun_out/c15_bench_n2.err
timeout 20 python scratch/show_bench.py gpurun_out/c15_bench_n2.json
