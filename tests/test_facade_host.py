"""CPU-side tests of the HemoCell C++ API surface (include/hemocell.h, hemocell_b200/host/facade.cpp):
the example case builds and links, host-side set-up works without a device, and the product fails
loudly (no CPU fallback) when no CUDA device is present."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _build():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    if not os.path.exists(os.path.join(ROOT, "hemocell_b200", "libhemocell_gpu.so")):
        G.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s"])


def test_host_facade_unit(tmp_path):
    _build()
    exe = tmp_path / "test_host_facade"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{ROOT}/include", f"-I{ROOT}/include/compat",
                           f"{ROOT}/tests/cpp/test_host_facade.cpp", "-o", str(exe), f"-L{ROOT}/hemocell_b200", "-lhemocell_gpu",
                           f"-Wl,-rpath,{ROOT}/hemocell_b200"])
    work = tmp_path / "work"; work.mkdir()
    r = subprocess.run([str(exe), str(work)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout and "FAIL" not in r.stdout


def _has_gpu():
    import ctypes as C
    try:
        rt = C.CDLL("libcudart.so.12"); n = C.c_int(0)
        return rt.cudaGetDeviceCount(C.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def test_example_fails_loudly_without_device(tmp_path):
    """no CPU fallback: the case file stops at the first device use with a clear message"""
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    _build()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import facade_cases as F
    F.write_shear_case(tmp_path, tmax=10, tmeas=10)
    r = subprocess.run([os.path.join(ROOT, "examples", "shear_cell", "shear_cell"), "config.xml"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
    assert not (tmp_path / "shear.log").exists() or (tmp_path / "shear.log").read_text() == ""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (authoring container only)")
@pytest.mark.parametrize("case", [
    "examples/oneCellShear/oneCellShear.cpp", "cases/performance_testing/performance_testing.cpp", "examples/cube/cube.cpp",
    "examples/stretchCell/stretchCell.cpp", "cases/stenosis/stenosis.cpp", "examples/simple/simple.cpp",
    "examples/parallelplanes/parallelplanes.cpp", "examples/flowaroundsphere/flowaroundsphere.cpp",
    "examples/microcontraction/microcontraction.cpp", "examples/capillary/wedge.cpp", "cases/atherosclerosis/atherosclerosis.cpp",
    "cases/cellCollision/cellCollision.cpp", "cases/kolmogorovFlow/kolmogorovFlow.cpp", "cases/microvessel_bended/microvessel_bended.cpp",
    "cases/stentflow/stentflow.cpp", "cases/unbounded/unbounded.cpp", "cases/vasoconstriction_pipe/vasoconstriction_pipe.cpp",
    "examples/pipeflow/pipeflow.cpp", "examples/parachuting/parachuting.cpp",
    # pre-inlet cases (hemo::PreInlet, Zou-He inlet / pressure outlet)
    "examples/pipeflow_with_preinlet/pipeflow_with_preinlet.cpp", "examples/curvedflow_with_preinlet/curvedflow_with_preinlet.cpp",
    "cases/AR2/AR2.cpp", "cases/AR2_pulsatile/AR2_pulsatile.cpp", "cases/AR2_stiff/AR2_stiff.cpp"])
def test_reference_case_files_compile_unmodified(case):
    """drop-in check: the reference's own case files compile against include/hemocell.h as they are"""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", f"-I{ROOT}/include", f"-I{ROOT}/include/compat", os.path.join(REF, case)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
