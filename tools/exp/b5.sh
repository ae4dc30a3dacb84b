for m in 2 3; do HCG_TAU1=0 HCG_K1_MINB=$m python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb=$m', round(d['value']), d['kernel_ms_per_step'].get('kernel:k_collide_stream'))"; done
