#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
exec < /dev/null
for c in "2.1,3.7,9.5,0.8,1.3" "2.1,3.7,8.8,0.8,1.3" "2.1,3.9,8.6,0.9,1.2"; do
  ZA_PROVER_COSTS=$c ZA_DEBUG_TIMELINE=1 timeout 100 python scratch/r2_prover_tl.py 8 2> gpurun_out/r2c39_tl8_$c.log | tail -1
  echo "costs $c: $(grep 'proof [0-9]*:' gpurun_out/r2c39_tl8_$c.log | tail -3 | tr '\n' ' ')"
  grep "collected" gpurun_out/r2c39_tl8_$c.log | tail -8 | sed 's/.*dev \([0-9]\).*collected \(.*\)/dev \1 \2/' | tr '\n' ' '; echo
done
