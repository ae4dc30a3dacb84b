"""dev: time the lattice kernel alone for each occupancy variant (HCG_K1_MINB)"""
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
    print(f"MINB={os.environ.get('HCG_K1_MINB')}: {ms/50:.4f} ms/step {n**3/(ms/50)/1e3:.0f} MLUPS {n**3*304/(ms/50)/1e6:.0f} GB/s")
else:
    for v in ("1", "2", "3", "4"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, HCG_K1_MINB=v))
