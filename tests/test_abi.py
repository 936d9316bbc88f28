"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("uvt.h", "uvt_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(uvt_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    # static inline helpers are not exported
    return names - {"uvt_voxel", "uvt_lcg_rand"}


def test_library_exports_every_declared_symbol(uvt):
    L = uvt._native.load()
    declared = _declared_symbols()
    assert len(declared) > 60
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # and the Python binding table covers the same set
    assert declared == set(uvt._native.SIGNATURES), declared ^ set(uvt._native.SIGNATURES)


def test_abi_version_and_struct_sizes(uvt):
    N = uvt._native
    assert N.load().uvt_abi_version() == 2
    assert N.CAMERA_DTYPE.itemsize == 96     # camera.zig:12-16 / std140
    assert N.CAMERA_DTYPE.fields["cam_mat"][1] == 16 and N.CAMERA_DTYPE.fields["fov"][1] == 80
    assert N.HIT_DTYPE.itemsize == 28
    assert ctypes.sizeof(N.Params) == 64
    p = N.Params()
    N.load().uvt_default_params(ctypes.byref(p))
    assert (p.map_dim, p.primary_max_steps, p.shadow_max_steps, p.edit_max_steps) == (512, 192, 48, 64)
    assert np.float32(p.epsilon) == np.float32(0.001)


def test_no_cpu_fallback_without_device(uvt):
    """Without a GPU uvt_create must fail loudly; nothing routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(uvt.UvtError) as e:
        uvt.Context(0)
    assert e.value.status == uvt._native.UVT_ERR_NO_DEVICE
    assert "no CPU path" in e.value.message


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "unnamed-voxel-tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f


def test_group_create_fails_loudly_without_a_gpu():
    import ctypes
    import importlib
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    uvt = importlib.import_module("unnamed-voxel-tracer_b200")
    L = uvt._native.load()
    devs = (ctypes.c_int * 2)(0, 1)
    h = ctypes.c_void_p()
    rc = L.uvt_group_create(None, devs, 2, ctypes.byref(h))
    assert rc == uvt._native.UVT_ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in L.uvt_group_last_error(None)
