#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 400 python -m pytest tests/test_prover_multi.py -k "matches_oracle or synthetic_key" -m gpu -x -q ) > gpurun_out/r2c40_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c40_pytest.log
grep -v "^\[za" gpurun_out/r2c40_pytest.log | tail -5
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c40_bench_n$N.json 2> gpurun_out/r2c40_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c40_bench_n$N.json | head -2
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c40_bench_n$N.json 2> gpurun_out/r2c40_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c40_bench_n$N.json | head -2
