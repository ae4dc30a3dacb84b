python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for o in 1 0; do HCG_SYNC_OVERLAP=$o python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 2952$o bench.py --gpus 2 --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['kernel_ms_per_step']; print('overlap=$o', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), {x:round(k[x],3) for x in ('syncEnvelopes','advanceParticles','interpolateFluidVelocity')})"; done
