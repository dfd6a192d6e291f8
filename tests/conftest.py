import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ms():
    import medseg_b200
    if not os.path.exists(medseg_b200.LIB_PATH):
        medseg_b200.build()
    return medseg_b200


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle on the plain-C oracle (oracle/c/medseg_oracle.c)."""
    import ctypes
    so = os.path.join(ROOT, "oracle", "_build", "libmedseg_oracle.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "c")], check=True, capture_output=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def stage_engine(ms):
    e = ms.Engine(None)
    yield e
    e.cleanup()


@pytest.fixture(scope="session")
def blob3(ms, tmp_path_factory):
    p = str(tmp_path_factory.mktemp("w") / "unet3.msegw")
    return ms.make_weight_blob(p, n_classes=3, seed=1234)


@pytest.fixture(scope="session")
def unet_engine(ms, blob3, tmp_path_factory):
    log_dir = str(tmp_path_factory.mktemp("log"))
    e = ms.Engine({"weights": blob3, "max_batch": 4}, log_dir)
    yield e
    e.cleanup()


@pytest.fixture(scope="session")
def torch_unet3(blob3):
    from medseg_b200 import weights as W
    from oracle.unet_torch import load_unet
    arch, w = W.load_blob(blob3)
    return load_unet(w, arch["n_classes"])


def contours_equal(a, b) -> bool:
    return len(a) == len(b) and all(np.asarray(x).shape == np.asarray(y).shape and (np.asarray(x) == np.asarray(y)).all()
                                    for x, y in zip(a, b))
