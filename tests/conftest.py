"""pytest configuration.  `-m "not gpu"` runs here (no GPU): the host logic, the g++ emulation of the device code
against the golden fixtures / the compiled reference, the restated oracle, the C-ABI export check.  `-m gpu` runs on
a B200: the parity tests proper, all through the C ABI (libpbrgpu.so)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import pbrlab_b200 as pb  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def built():
    """Native libraries present (built in-tree by `make`)."""
    need = [pb.GPU_LIB, pb.HOST_LIB, os.path.join(HERE, "host_emul", "libpbr_emul.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", ROOT, "-j8", "gpu", "host", "emul"], check=True)
    return True


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref) or None when it did not travel / was not built."""
    import refbind
    if not refbind.available():
        return None
    return refbind.RefLib()


@pytest.fixture(scope="session")
def cornell_host(built):
    """pbrlab::Scene of the bundled Cornell scene, host-side commit only (no device)."""
    return pb.Scene([scenes.cornell()], commit_to_device=False)


@pytest.fixture(scope="session")
def cornell_emul(cornell_host):
    import emulbind
    return emulbind.Emul(cornell_host.flat())


@pytest.fixture(scope="session")
def hair_file():
    path = os.path.join(scenes.CACHE, "golden_hair.hair")
    os.makedirs(scenes.CACHE, exist_ok=True)
    scenes.write_cyhair(path, n_strands=400, n_points=9, center=(-2.5, 6.0, 0.0), radius=1.0, length=2.0,
                        thickness=0.02, seed=99)
    return path


@pytest.fixture(scope="session")
def hair_host(built, hair_file):
    return pb.Scene([scenes.cornell(), hair_file], commit_to_device=False)


@pytest.fixture(scope="session")
def hair_emul(hair_host):
    import emulbind
    return emulbind.Emul(hair_host.flat())


@pytest.fixture(scope="session")
def cornell_gpu(built):
    s = pb.Scene([scenes.cornell()])
    return s, s.context()


@pytest.fixture(scope="session")
def hair_gpu(built, hair_file):
    s = pb.Scene([scenes.cornell(), hair_file])
    return s, s.context()
