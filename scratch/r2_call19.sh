#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
nvidia-smi -L | wc -l
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c19_bench_n$N.json 2> gpurun_out/r2c19_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c19_bench_n$N.json
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c19_bench_n$N.json 2> gpurun_out/r2c19_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c19_bench_n$N.json
( time timeout 300 python -m pytest tests/test_prover_multi.py -k "matches_oracle or synthetic_key" -m gpu -x -q ) > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c19_pytest.log
grep -v "^\[za" gpurun_out/r2c19_pytest.log | tail -5
ZA_DEBUG_TIMELINE=1 timeout 120 python scratch/r2_prover_tl.py 8 2> gpurun_out/r2c19_tl8.log | tail -1
grep "proof [0-9]*:" gpurun_out/r2c19_tl8.log | tail -3
ZA_SWEEP_SIZES=16,18,20,22,24,26 timeout 400 python scratch/sweep_config5_multi.py 26 2,4,8 16 > gpurun_out/r2c19_sweep_multi.md 2> gpurun_out/r2c19_sweep_multi.err; tail -20 gpurun_out/r2c19_sweep_multi.md; tail -3 gpurun_out/r2c19_sweep_multi.err
