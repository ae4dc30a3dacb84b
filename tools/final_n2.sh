#!/bin/bash
# two-GPU confirmation of a build: multi-rank + pre-inlet tests, then the benchmark line at N = 2
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_multi.py tests/test_gpu_zz_preinlet.py -q 2>&1 | tail -4 > gpurun_out/r2z_gpu_tests_2gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2z_bench_n2_c1.json 2> gpurun_out/r2z_bench_n2.err
tail -2 gpurun_out/r2z_gpu_tests_2gpu.txt; tail -c 300 gpurun_out/r2z_bench_n2.err
