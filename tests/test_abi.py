"""CPU: the C-ABI library builds/loads and exports exactly what include/pv_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    names = _header_functions()
    assert "pv_conv3d_fwd" in names and "pv_clip_transform_fwd" in names and len(names) >= 15


def test_library_loads_and_exports_every_declared_symbol():
    from pytorchvideo_b200 import _lib
    lib = _lib.load()
    for name in _header_functions():
        assert hasattr(lib, name), "libpvb200.so does not export %s" % name
    assert lib.pv_abi_version() == 1
    # the ctypes table binds exactly the declared functions
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_exports_are_c_linkage():
    from pytorchvideo_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for name in _header_functions():
        assert name in exported


def test_no_device_is_reported_not_faked():
    import torch
    from pytorchvideo_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.require_device()


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (compile a tiny probe with gcc)."""
    from pytorchvideo_b200 import _lib
    probe = r'''
    #include <stdio.h>
    #include "pv_b200.h"
    int main(){ printf("%zu %zu %zu %zu %zu\n", sizeof(pv_clip_transform_desc), sizeof(pv_conv3d_desc),
                       sizeof(pv_pool3d_desc), sizeof(pv_attention_desc), sizeof(pv_clip_batch_desc)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(probe)
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.ClipTransformDesc), ctypes.sizeof(_lib.Conv3dDesc),
                     ctypes.sizeof(_lib.Pool3dDesc), ctypes.sizeof(_lib.AttentionDesc), ctypes.sizeof(_lib.ClipBatchDesc)]


def test_product_has_no_cpu_path():
    import torch
    import pytorchvideo_b200.models.hub as H
    from pytorchvideo_b200.transforms import FusedClipTransform
    m = H.x3d_xs().eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 4, 160, 160))
    with pytest.raises(RuntimeError):
        FusedClipTransform(4)(torch.zeros(3, 8, 16, 16, dtype=torch.uint8))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "pytorchvideo_b200")):
        for f in fs:
            if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(dp, f)).read(), re.M):
                bad.append(f)
    assert not bad
