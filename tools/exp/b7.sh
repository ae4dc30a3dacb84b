for pad in 0 40; do HCG_MECH_SMEM_PAD=$pad python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('pad=$pad', round(d['value']), 'mech per call ms', k['applyConstitutiveModel']*20)"; done
