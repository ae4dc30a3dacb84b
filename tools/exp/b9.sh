python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('main', round(d['value']), d['ms_per_step'], 'mech per call ms', k['applyConstitutiveModel']*20)"
