"""CPU tests of the multi-GPU host logic: slab membership (hch_slab_membership) is symmetric
between neighbouring ranks, and a world_size-2 gloo run agrees on the shared lists."""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cells(seed, n, nx):
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-10, nx + 10, n)
    return lo, lo + rng.uniform(4, 16, n)


def test_membership_symmetry_and_coverage():
    from hemocell_b200 import lib as H
    nx, R, M = 256, 4, 4.0
    nxl = nx // R
    lo, hi = _cells(0, 4000, nx)
    res = [H.slab_membership(lo, hi, nx, True, nxl, r, R, M) for r in range(R)]
    held = np.array([r[0] for r in res])
    assert held.any(0).all()                                  # every cell is held somewhere
    for r in range(R):
        right = (r + 1) % R
        # what r shares through its right face is exactly what its right neighbour shares through its left face
        assert np.array_equal(res[r][2], res[right][1])
        assert np.all(held[r][res[r][2]]) and np.all(held[right][res[r][2]])
    # a cell wholly inside a slab, away from the faces, is held by exactly one rank and shared by none
    inside = (lo % nx > 10) & (hi % nx < nxl - 10) & (hi - lo < 16) & (lo > 0) & (hi < nx)
    assert np.all(held[:, inside].sum(0) == 1)
    # non-periodic ends
    r0 = H.slab_membership(lo, hi, nx, False, nxl, 0, R, M)
    assert not r0[1].any()
    one = H.slab_membership(lo, hi, nx, True, nx, 0, 1, M)
    assert one[0].all() and not one[1].any() and not one[2].any()


def test_two_rank_gloo_agreement(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch, torch.distributed as dist
from hemocell_b200 import lib as H
dist.init_process_group("gloo")
r, R = dist.get_rank(), dist.get_world_size()
nx = 192; nxl = nx // R
rng = np.random.default_rng(5)
lo = rng.uniform(0, nx, 3000); hi = lo + rng.uniform(4, 16, 3000)
held, sl, sr = H.slab_membership(lo, hi, nx, True, nxl, r, R, 4.0)
mine = torch.tensor(np.stack([sl, sr]).astype(np.uint8))
both = [torch.zeros_like(mine) for _ in range(R)]
dist.all_gather(both, mine)
other = both[1 - r]
# with two ranks and periodic x the neighbour is the same rank through both faces
assert torch.equal(mine[0], other[1]) and torch.equal(mine[1], other[0]), "shared lists disagree"
n = torch.tensor([int(held.sum())]); dist.all_reduce(n)
assert n.item() >= 3000
print("rank", r, "ok", int(sl.sum()), int(sr.sum()))
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2
