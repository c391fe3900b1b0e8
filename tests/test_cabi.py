"""The drop-in boundary: libpbrgpu.so loads and exports every symbol include/pbrgpu.h declares; without a GPU the
backend refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import pbrlab_b200 as pb
from conftest import ROOT, _have_gpu


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pbrgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbrgpu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = C.CDLL(pb.GPU_LIB)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libpbrgpu.so does not export " + n


def test_struct_layouts(built):
    assert C.sizeof(pb.Material) == 112
    assert pb.RAY_DTYPE.itemsize == 32 and pb.HIT_DTYPE.itemsize == 36


@pytest.mark.skipif(_have_gpu(), reason="checks the GPU-less failure mode")
def test_no_cpu_fallback(built):
    lib = pb.gpu_lib()
    assert not lib.pbrgpu_create(None, 0)
    msg = lib.pbrgpu_last_error(None).decode()
    assert "no CPU fallback" in msg
    with pytest.raises(RuntimeError):
        pb.Context()
    from pbrlab_b200 import scenes
    with pytest.raises(RuntimeError):
        pb.Scene([scenes.cornell()])   # CommitScene() throws: no B200 backend
