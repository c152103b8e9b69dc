#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
( time timeout 600 python -m pytest tests/test_prover_multi.py tests/test_multi_gpu.py tests/test_gpu_parity.py -k "prover or two_gpu or staged or contexts" -m gpu -x -q ) > gpurun_out/r2c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c16_pytest.log
grep -v "^\[za" gpurun_out/r2c16_pytest.log | tail -8
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-sub > gpurun_out/r2c16_bench_n$N.json 2> gpurun_out/r2c16_bench_n$N.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c16_bench_n$N.json
ZA_PROVER_PLAN=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2c16_bench_n${N}_slices.json 2> gpurun_out/r2c16_bench_n${N}_slices.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c16_bench_n${N}_slices.json
ZA_PROVER_H_COPY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/r2c16_bench_n${N}_hcopy.json 2> gpurun_out/r2c16_bench_n${N}_hcopy.err
timeout 20 python scratch/show_bench.py gpurun_out/r2c16_bench_n${N}_hcopy.json
ZA_DEBUG_TIMELINE=1 timeout 120 python scratch/r2_prover_tl.py 2 2> gpurun_out/r2c16_tl2.log | tail -2
grep -v "timeline\] [LA] " gpurun_out/r2c16_tl2.log | tail -12
