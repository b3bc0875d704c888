import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): builds oracle/_build/libsrk_oracle.so on first use."""
    import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def engine():
    import surikatoko_b200 as sb
    eng = sb.Engine(0)
    yield eng
    eng.close()


def to_problem(pr):
    """oracle_lib.Problem -> surikatoko_b200.BAProblem (copies)."""
    import surikatoko_b200 as sb
    return sb.BAProblem(pr.obs_cam.copy(), pr.obs_point.copy(), pr.obs_xy.copy(), pr.points.copy(), pr.cams.copy(), pr.K.copy(),
                        pr.shared_K, pr.f0)


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 0.0
    return d / s if s > 0 else d
