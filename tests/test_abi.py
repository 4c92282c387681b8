"""CPU: the C-ABI library builds, loads and exports every symbol include/salp_b200.h declares;
host-side argument checking works without a GPU; creating a simulator without a GPU fails
loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from grasp_lab_salp_b200 import _lib, default_params
from grasp_lab_salp_b200.params import SalpParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "salp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(salp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"libsalp_b200.so does not export {name}"
        assert name in _lib.PROTOTYPES, f"ctypes binding misses {name}"
    assert b"sm_100a" in lib.salp_build_info()


def test_params_struct_layout_matches_header():
    lib = _lib.load()
    p = SalpParams()
    assert lib.salp_default_params(C.byref(p)) == 0
    q = default_params()
    for name, _ in SalpParams._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if hasattr(a, "__len__"):
            np.testing.assert_allclose(list(a), list(b), rtol=1e-9, err_msg=name)
        else:
            assert a == pytest.approx(b, rel=1e-12), name
    # the C default fit and np.polyfit agree to rounding; the Python host always uploads np.polyfit's bits
    np.testing.assert_allclose(list(p.refill_poly), [-500.0, 87.0, -0.45], rtol=1e-9)
    np.testing.assert_allclose(list(p.jet_poly), [-250.0, 25.5, -0.125], rtol=1e-9)


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    p = default_params()
    assert lib.salp_create(None, 4, 0, 0, 0, C.byref(h)) == _lib.ERR_INVALID
    assert lib.salp_create(C.byref(p), 0, 0, 0, 0, C.byref(h)) == _lib.ERR_INVALID
    bad = p.copy()
    bad.num_obstacles = 99
    assert lib.salp_create(C.byref(bad), 4, 0, 0, 0, C.byref(h)) == _lib.ERR_INVALID
    assert b"num_obstacles" in lib.salp_last_error(None)
    assert lib.salp_step(None, None, 0, None) == _lib.ERR_INVALID
    assert lib.salp_destroy(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from grasp_lab_salp_b200 import SalpBatch, SalpError
    with pytest.raises(SalpError) as e:
        SalpBatch(8)
    assert e.value.code == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/salp_b200.h must compile as C99 (no C++ in the
    signatures, every type it uses declared)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "salp_b200.h"\nint main(void) { return (int)sizeof(SalpParams) == 0 || (int)sizeof(SalpStepIO) == 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                        str(src), "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
