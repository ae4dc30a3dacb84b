#!/bin/bash
# IBM kernels (spread, interpolate + advance): parity, timing, one ncu capture of each
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -x 2>&1 | tail -3
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
python tools/quick_variants.py HCG_SPREAD_BULK=1 | tee gpurun_out/r2y3_ibm_variants.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_spread_sorted|k_interp_advance" -c 2 -f -o gpurun_out/r2y3_ibm \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2y_ncu.log 2>&1
tail -2 gpurun_out/r2y_ncu.log
