#!/bin/bash
# GPU call 1 of round 2: parity of the new field layer on the device, variant timings, proof timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -3 gpurun_out/r2c1_pytest.log
for v in main kara r192; do
  if [ $v = main ]; then unset ZA_B200_SO; else export ZA_B200_SO=$PWD/za_b200/variants/libza_b200_$v.so; fi
  timeout 600 python scratch/r2_variant_time.py $v >> gpurun_out/r2c1_variants.log 2>&1
done
unset ZA_B200_SO
cat gpurun_out/r2c1_variants.log | grep "^\["
ZA_DEBUG_TIMELINE=1 timeout 300 python scratch/dbg_prove.py > gpurun_out/r2c1_timeline.log 2>&1
tail -30 gpurun_out/r2c1_timeline.log
