for lg in 16 18 20 22; do for c in 12 13 14 15 16 17 18 19 20; do
  if [ $c -le $((lg+1)) ]; then ZA_MSM_TABLE=$c python scratch/sweep_c.py $lg 1 2>/dev/null | tail -1; fi
done; done
