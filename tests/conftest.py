import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return ge.load_package()


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def O():
    from oracle import oracle

    return oracle


@pytest.fixture(scope="session")
def orc(O):
    """The C restatement (oracle/libmbavo_oracle.so); built on demand."""
    return O.OracleLib()


@pytest.fixture(scope="session")
def ref(O):
    """oracle/_ref: the reference's own header arithmetic.  Present where it was built from /root/reference (it also
    travels to the GPU box as a binary); otherwise the tests that need it are skipped and tests/golden covers them."""
    if not O.RefLib.available():
        pytest.skip("oracle/_ref/libmbavo_ref.so not built (no /root/reference on this machine)")
    return O.RefLib()


@pytest.fixture(scope="session")
def cuda_lib(pkg):
    """The product library.  Loading it needs no GPU; it must exist (there is no fallback)."""
    return pkg.load_library()


def reference_test_spline(n_knots=7):
    """create_spline of the reference test (test/test_blur_aware_tracker_modules.cpp:24-67): roll/pitch/yaw ramps,
    translation (5i, 5i, 0).  Transformation::setRollPitchYaw builds Rz(yaw) Ry(pitch) Rx(roll)."""
    import numpy as np

    rpy = [(0.01, 0.01, 0.002), (0.02, 0.015, 0.0015), (0.03, 0.02, 0.001), (0.04, 0.025, 0.0005), (0.05, 0.03, 0.0),
           (0.05, 0.035, -0.0005), (0.07, 0.04, -0.001)][:n_knots]
    kt = np.array([[5.0 * i, 5.0 * i, 0.0] for i in range(n_knots)])
    kR = []
    for r, p, y in rpy:
        r, p, y = r * np.pi, p * np.pi, y * np.pi
        cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
        kR.append([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                   cr * cp * cy + sr * sp * sy])
    return kt, np.array(kR)
