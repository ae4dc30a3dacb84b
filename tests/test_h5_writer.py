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


def _libhdf5():
    """h5py, else libhdf5 through ctypes, else None (neither is in the authoring image; the test runs wherever one exists)"""
    try:
        import h5py                                   # noqa: F401
        return "h5py"
    except Exception:
        pass
    import ctypes.util
    for name in ("hdf5", "hdf5_serial"):
        if ctypes.util.find_library(name):
            return ctypes.util.find_library(name)
    return None


@pytest.mark.skipif(_libhdf5() is None, reason="neither h5py nor libhdf5 on this machine")
def test_files_open_with_the_real_hdf5_library(tmp_path):
    """the consumers of the reference's output (scripts/*XMF.py, ParaView) use libhdf5: wherever h5py or libhdf5 exists, the
    files of our writer must open with it and show the layout of io/ParticleHdf5IO.cpp:60-194 (dataset names, shapes, element
    types, root attributes)"""
    rng = np.random.default_rng(5)
    n = 642*3
    arrays = {"Position": rng.normal(size=(n, 3)).astype(np.float32), "Total force": rng.normal(size=(n, 3)).astype(np.float32),
              "Cell Id": rng.integers(0, 99, size=(n, 1)).astype(np.float32), "Triangles": rng.integers(0, n, size=(1280*3, 3)).astype(np.int32)}
    chunks = {k: (min(1000, v.shape[0]), v.shape[1]) for k, v in arrays.items()}
    attrs = {"dx": np.array([5e-7]), "dt": np.array([1e-7]), "iteration": np.array([1200], dtype=np.int64),
             "processorId": np.array([0], dtype=np.int32), "numberOfParticles": np.array([n], dtype=np.int64)}
    p = tmp_path / "RBC.000000001200.p.0.h5"
    _write(p, 7, arrays, attrs, chunks)
    how = _libhdf5()
    if how == "h5py":
        import h5py
        with h5py.File(p, "r") as f:
            assert sorted(f.keys()) == sorted(arrays)
            for k, v in arrays.items():
                assert f[k].shape == v.shape and f[k].dtype == v.dtype and f[k].chunks == chunks[k] and f[k].compression == "gzip"
                np.testing.assert_array_equal(f[k][...], v)
            for k, v in attrs.items():
                np.testing.assert_array_equal(np.asarray(f.attrs[k]).reshape(-1), v)
    else:
        import ctypes as C
        h5 = C.CDLL(how)
        h5.H5open()
        h5.H5Fopen.restype = C.c_int64; h5.H5Dopen2.restype = C.c_int64; h5.H5Dget_space.restype = C.c_int64
        fid = h5.H5Fopen(str(p).encode(), C.c_uint(0), C.c_int64(0))
        assert fid >= 0, "H5Fopen rejected the file"
        for k, v in arrays.items():
            did = h5.H5Dopen2(C.c_int64(fid), k.encode(), C.c_int64(0))
            assert did >= 0, k
            sid = h5.H5Dget_space(C.c_int64(did))
            dims = (C.c_uint64 * 8)()
            rank = h5.H5Sget_simple_extent_dims(C.c_int64(sid), dims, None)
            assert tuple(dims[:rank]) == v.shape
            h5.H5Sclose(C.c_int64(sid)); h5.H5Dclose(C.c_int64(did))
        h5.H5Fclose(C.c_int64(fid))
