"""HDF5 container writer (hemocell_b200/host/hemo_h5.cpp, behind HemoCell::writeOutput) read back with the
independent format-level reader tests/h5mini.py: structure, types, shapes, attributes, chunk index, deflate."""
import numpy as np
import pytest

import h5mini
from hemocell_b200 import lib as H


def _write(path, level, arrays, attrs, chunks):
    w = H.H5Writer(path, level)
    for k, v in attrs.items():
        w.attribute(k, v)
    for k, v in arrays.items():
        w.dataset(k, v, chunks.get(k))
    w.close()


@pytest.mark.parametrize("level", [7, -1])
def test_particle_like_file_round_trip(tmp_path, level):
    rng = np.random.default_rng(3)
    n = 642*5 + 17                                   # not a multiple of the 1000-row chunk
    arrays = {"Position": rng.normal(size=(n, 3)).astype(np.float32),
              "Total force": rng.normal(size=(n, 3)).astype(np.float32),
              "Cell Id": rng.integers(0, 99, size=(n, 1)).astype(np.float32),
              "Triangles": rng.integers(0, n, size=(1280*5, 3)).astype(np.int32)}
    chunks = {k: (min(1000, v.shape[0]), v.shape[1]) for k, v in arrays.items()}
    attrs = {"dx": np.array([5e-7]), "dt": np.array([1e-7]), "iteration": np.array([1200], dtype=np.int64),
             "processorId": np.array([0], dtype=np.int32), "numberOfParticles": np.array([n], dtype=np.int64)}
    p = tmp_path / "RBC.000000001200.p.0.h5"
    _write(p, level, arrays, attrs, chunks)
    f = h5mini.File(p)
    assert sorted(f.datasets) == sorted(arrays)
    for k, v in arrays.items():
        d = f.datasets[k]
        assert d["dtype"] == v.dtype and d["shape"] == v.shape
        assert d["layout"] == ("chunked" if level >= 0 else "contiguous")
        if level >= 0:
            assert d["deflate"] == 7 and d["chunk"] == chunks[k] and d["nchunks"] == -(-v.shape[0]//chunks[k][0])
        np.testing.assert_array_equal(d["data"], v)
    for k, v in attrs.items():
        assert f.attrs[k].dtype == v.dtype
        np.testing.assert_array_equal(f.attrs[k], v)


def test_fluid_like_4d_dataset_and_many_chunks(tmp_path):
    rng = np.random.default_rng(4)
    vel = rng.normal(size=(12, 10, 9, 3)).astype(np.float32)
    big = np.arange(70001*2, dtype=np.float64).reshape(-1, 2)        # 7001 chunks of 10 rows: a 3-level chunk B-tree
    p = tmp_path / "Fluid.h5"
    _write(p, 1, {"Velocity": vel, "big": big, "empty": np.zeros((0, 3), np.float32)},
           {"subdomainSize": np.array([12, 10, 9], dtype=np.int32), "relativePosition": np.array([-1.5, -1.5, -1.5], dtype=np.float32)},
           {"Velocity": (5, 10, 4, 3), "big": (10, 2), "empty": (1, 3)})
    f = h5mini.File(p)
    np.testing.assert_array_equal(f["Velocity"], vel)
    assert f.datasets["Velocity"]["nchunks"] == 3*1*3*1
    np.testing.assert_array_equal(f["big"], big)
    assert f.datasets["big"]["nchunks"] == 7001
    assert f["empty"].shape == (0, 3)
    np.testing.assert_array_equal(f.attrs["subdomainSize"], [12, 10, 9])


def test_many_datasets_single_symbol_node(tmp_path):
    arrays = {"d%02d" % i: np.full((3, 2), i, dtype=np.int64) for i in range(23)}
    p = tmp_path / "many.h5"
    _write(p, -1, arrays, {}, {})
    f = h5mini.File(p)
    assert f.leaf_k >= 12
    for k, v in arrays.items():
        np.testing.assert_array_equal(f[k], v)
