#!/bin/bash
# mechanics kernel: parity, timing at ring loop unrolled by 1 / 2, one full ncu capture of the fastest
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -q -k "mechanics or iterate" 2>&1 | tail -3
python -m pytest tests/test_gpu_configs.py -q -x 2>&1 | tail -3
python tools/quick_variants.py HCG_MECH_UNROLL=1 HCG_MECH_UNROLL=2 | tee gpurun_out/r2x_mech_variants.txt
best=$(python - <<'P'
import re
b=None
for l in open("gpurun_out/r2x_mech_variants.txt"):
    m=re.search(r"HCG_MECH_UNROLL=(\d+).*'applyConstitutiveModel': ([0-9.]+)", l)
    if m and (b is None or float(m.group(2))<b[1]): b=(m.group(1),float(m.group(2)))
print(b[0] if b else 1)
P
)
echo "best $best"
HCG_MECH_UNROLL=$best timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mechanics -c 1 -f -o gpurun_out/r2x_k_mechanics \
  python bench.py --steps 25 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2x_ncu.log 2>&1
tail -2 gpurun_out/r2x_ncu.log
