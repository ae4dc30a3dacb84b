#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root): verifies and measures what round 1 left unmeasured.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_gpu.sh'
# 1. the gated tests (moment-only update at tau = 1, curvedflow_with_preinlet smoke run)
# 2. bench with stored populations vs the moment-only update (same box, back to back)
# 3. launch list + one --set full capture of k_moment_step
set -u
mkdir -p gpurun_out
HCG_TEST_MOMENT_ONLY=1 timeout 300 python -m pytest tests/test_gpu_zzz_unverified.py -q 2>&1 | tail -30 > gpurun_out/nr_tests.log
timeout 200 python bench.py --steps 100 --warmup 10 > gpurun_out/nr_bench_pops.json 2> gpurun_out/nr_bench_pops.err
HCG_MOMENT_ONLY=1 timeout 200 python bench.py --steps 100 --warmup 10 > gpurun_out/nr_bench_moment_only.json 2> gpurun_out/nr_bench_moment_only.err
HCG_MOMENT_ONLY=1 HCG_MOMENT_STATE=vel timeout 200 python bench.py --steps 100 --warmup 10 > gpurun_out/nr_bench_moment_only_vel.json 2>> gpurun_out/nr_bench_moment_only.err
HCG_MOMENT_ONLY=1 timeout 200 python bench.py --steps 100 --warmup 10 --cadence 5 > gpurun_out/nr_bench_moment_only_c5.json 2>> gpurun_out/nr_bench_moment_only.err
HCG_MOMENT_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/nr_launches.csv \
  python bench.py --steps 4 --warmup 3 > gpurun_out/nr_ncu_a.log 2>&1
HCG_MOMENT_ONLY=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_moment_step -c 2 -o gpurun_out/nr_k_moment_step \
  python bench.py --steps 4 --warmup 3 > gpurun_out/nr_ncu_b.log 2>&1
# (multi-GPU, separate calls: gpurun --gpus 2 -- 'HCG_TEST_MOMENT_ONLY=1 python -m pytest tests/test_gpu_zzz_unverified.py -q -k two_gpu';
#  and gpurun --gpus 2 -- 'HCG_MOMENT_ONLY=2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10')
tail -3 gpurun_out/nr_tests.log; cat gpurun_out/nr_bench_pops.json gpurun_out/nr_bench_moment_only.json | cut -c1-400
