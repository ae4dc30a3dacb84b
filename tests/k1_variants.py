"""dev: time the lattice kernels alone for each variant (env HCG_K1_ROWS / HCG_K1_STAGES / HCG_K1_CTAS)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    from hemocell_b200 import lib as H
    n = 256
    ctx = H.Context(n, n, n, (1, 1, 1), 1.0)
    ctx.set_flags(np.zeros(n**3, dtype=np.uint8))
    ctx.set_body_force((1e-7, 1e-7, 1e-7))
    ctx.fluid_warmup(10)
    ms = ctx.iterate_timed(50)
    tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("HCG_"))
    print(f"{tag}: {ms/50:.4f} ms/step {n**3/(ms/50)/1e3:.0f} MLUPS {n**3*304/(ms/50)/1e6:.0f} GB/s", flush=True)
else:
    combos = [dict(HCG_K1_ROWS="0")] + [dict(HCG_K1_NT=str(t), HCG_K1_STAGES=str(s), HCG_K1_CTAS=str(c), HCG_K1_INTERLEAVE=str(i)) for (t, s, c, i) in
              ((288, 4, 1, 0), (288, 4, 1, 1), (288, 2, 2, 1), (288, 3, 1, 1), (544, 2, 1, 1))]
    for v in combos:
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, **v), timeout=120)
