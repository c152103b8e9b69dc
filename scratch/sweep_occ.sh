cp za_b200/libza_b200.so /tmp/orig.so
for v in 4_3 5_3 6_3; do cp scratch/libza_$v.so za_b200/libza_b200.so; echo "== $v"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],3), d['kernel_ms_per_step'])"; done
cp /tmp/orig.so za_b200/libza_b200.so
