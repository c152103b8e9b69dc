for sl in 1 2 3 4; do for lm in 16 20; do
 echo "SEG_LOG=$sl log_m=$lm"; ZA_MSM_SEG_LOG=$sl python bench.py --steps 5 --warmup 3 --log-m $lm --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],3), round(d['e2e']['value'],3), d['gpu_launches'])"
done; done
