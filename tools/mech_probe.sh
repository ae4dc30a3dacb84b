#!/bin/bash
# mechanics kernel: parity, timing (one call in 30 steps: ms per call = 30 x the applyConstitutiveModel figure), one full ncu capture
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py -q -k "mechanics or iterate" 2>&1 | tail -3
python -m pytest tests/test_gpu_configs.py -q -x 2>&1 | tail -3
python tools/quick_variants.py "" | tee gpurun_out/mech_timing.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mechanics -c 1 -f -o gpurun_out/k_mechanics \
  python bench.py --steps 25 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/mech_ncu.log 2>&1
tail -2 gpurun_out/mech_ncu.log
