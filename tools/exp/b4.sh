python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('main', round(d['value']), d['ms_per_step'], d['kernel_ms_per_step'].get('kernel:k_collide_tau1'), d['kernel_ms_per_step'].get('kernel:k_moments'))"
