"""CPU tests of the C-ABI library and the host-side solver logic (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg, cuda_lib):
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mbavo.h")).read()
    declared = set(re.findall(r"\b(mbavo_[a-z_0-9]+)\s*\(", header))
    declared -= {"mbavo_ctx"}
    assert declared == set(pkg.EXPORTED_SYMBOLS), declared ^ set(pkg.EXPORTED_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(cuda_lib, name), name
    assert cuda_lib.mbavo_version() == 100
    assert cuda_lib.mbavo_packed_len(2) == 91 and cuda_lib.mbavo_packed_len(4) == 325  # (6k+1)(6k+2)/2, …cost.cu:209-210


def test_no_cpu_fallback_without_gpu(pkg):
    """Without a CUDA device the product fails loudly (MBAVO_ECUDA); it never computes on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MbavoError) as e:
        pkg.Context(pkg.Limits())
    assert "mbavo error -2" in str(e.value) or "mbavo error" in str(e.value)


def test_trust_region_step_matches_numpy(pkg, cuda_lib, O):
    """mbavo_trust_region_step (tracker.cpp:799-831 + solve_normal_equation.h) against the numpy restatement, both
    solver branches, including the in-place compounding damping."""
    rng = np.random.default_rng(4)
    for dim in (12, 18, 30, 42, 48):
        M = rng.normal(size=(dim, dim))
        scale = np.concatenate([np.full(dim // 2, 1.0), np.full(dim - dim // 2, 300.0)])  # t / w blocks differ by orders
        A = (M @ M.T + dim * np.eye(dim)) * np.outer(scale, scale)
        g = rng.normal(size=dim) * scale
        for solver, name in ((0, "SVD_JACOBI"), (1, "LDLT")):
            H1, H2 = A.copy(), A.copy()
            for radius in (1e4, 5e3):  # a rejected step re-damps the already damped matrix
                step = np.zeros(dim)
                model = C.c_double(0)
                rc = cuda_lib.mbavo_trust_region_step(H1.ctypes.data_as(C.POINTER(C.c_double)),
                                                      g.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(dim),
                                                      C.c_double(radius), C.c_int(solver),
                                                      step.ctypes.data_as(C.POINTER(C.c_double)), C.byref(model))
                assert rc == 0
                step_ref, model_ref = O.trust_region_step(H2, g, radius, name)
                assert np.allclose(H1, H2, rtol=0, atol=0)
                assert np.linalg.norm(step - step_ref) <= 1e-9 * np.linalg.norm(step_ref)
                assert abs(model.value - model_ref) <= 1e-9 * abs(model_ref)
                assert np.abs(H1 @ step + g).max() <= 1e-7 * np.abs(g).max()


def test_spline_plus_matches_oracle(cuda_lib, O):
    """mbavo_spline_plus (Spline.h:307-330): t += dt, R = R Exp(dw), no re-normalisation."""
    rng = np.random.default_rng(5)
    for n in (2, 3, 7):
        kt = rng.normal(size=(n, 3))
        kR = rng.normal(size=(n, 4))
        kR /= np.linalg.norm(kR, axis=1, keepdims=True)
        for scale in (1e-12, 1e-3, 0.3):
            step = rng.normal(size=6 * n) * scale
            ct, cR = np.zeros_like(kt), np.zeros_like(kR)
            dp = C.POINTER(C.c_double)
            assert cuda_lib.mbavo_spline_plus(C.c_int(n), kt.ctypes.data_as(dp), kR.ctypes.data_as(dp), step.ctypes.data_as(dp),
                                              ct.ctypes.data_as(dp), cR.ctypes.data_as(dp)) == 0
            rt, rR = O.plus(kt, kR, step)
            assert np.abs(ct - rt).max() <= 1e-15 and np.abs(cR - rR).max() <= 1e-15


def test_unpack_layout(cuda_lib):
    """mbavo_unpack == merge_hessian_gradient_cost.cpp:25-86: packed [cost, g, triu(H)] of a knot window scattered to
    global indices 3*(kmin+j) (t) and 3*(n+kmin+j) (w)."""
    rng = np.random.default_rng(6)
    dp = C.POINTER(C.c_double)
    for n, kmin, NK in ((2, 0, 2), (3, 0, 3), (5, 1, 3), (7, 3, 4)):
        d = 6 * NK
        A = rng.normal(size=(d + 1, d + 1))
        A = A + A.T
        packed = np.array([A[a, b] for a in range(d + 1) for b in range(a, d + 1)])
        cost = C.c_double(0)
        H, g = np.full((6 * n, 6 * n), np.nan), np.full(6 * n, np.nan)
        assert cuda_lib.mbavo_unpack(packed.ctypes.data_as(dp), C.c_int(kmin), C.c_int(NK), C.c_int(n), C.byref(cost),
                                     H.ctypes.data_as(dp), g.ctypes.data_as(dp)) == 0
        idx = [3 * kmin + j for j in range(3 * NK)] + [3 * (n + kmin) + j for j in range(3 * NK)]
        Hx, gx = np.zeros((6 * n, 6 * n)), np.zeros(6 * n)
        Hx[np.ix_(idx, idx)] = A[1:, 1:]
        gx[idx] = A[0, 1:]
        assert cost.value == A[0, 0] and np.array_equal(H, Hx) and np.array_equal(g, gx)
    # window outside the knots is refused
    assert cuda_lib.mbavo_unpack(packed.ctypes.data_as(dp), C.c_int(5), C.c_int(4), C.c_int(7), C.byref(cost), None, None) != 0


def test_lm_strategy_and_step_evaluator_restatement(O):
    """levenberg_marquardt_strategy.cpp:21-39 and trust_region_step_evaluator.cpp:45-126 (numpy restatement used by the
    oracle's LM loop): radius schedule and non-monotonic step quality."""
    lm = O.LMStrategy()
    assert lm.radius == 1e4
    lm.rejected()
    assert lm.radius == 5e3 and lm.decrease == 4.0
    lm.rejected()
    assert lm.radius == 1250.0 and lm.decrease == 8.0
    lm.accepted(1.0)                      # 1 - (2q-1)^3 = 0 -> clamp to 1/3 -> radius * 3
    assert lm.radius == 3750.0 and lm.decrease == 2.0
    lm.accepted(0.5)                      # factor 1
    assert lm.radius == 3750.0
    for _ in range(20):
        lm.rejected()
    assert lm.radius == 10.0              # min radius
    ev = O.StepEvaluator(5)
    ev.reset(10.0)
    assert ev.quality(9.0, 2.0) == 0.5
    ev.accepted(9.0, 2.0)
    assert ev.minimum == 9.0 and ev.reference == 10.0 and ev.acc_ref == 2.0
    assert ev.quality(8.5, 1.0) == max(0.5, 1.5 / 3.0)


def test_outlier_detection_restatement(O):
    """detectOutliersAndUploadToGpu (tracker.cpp:639-699): statistics exclude costs < 1e-8, the test does not; sticky flags."""
    c = np.array([1.0, 1.1, 0.9, 1.0, 0.0, 10.0, 1.05, 0.95])
    flags = np.zeros(8, dtype=np.uint8)
    n = O.detect_outliers(c, flags, 2.0)
    sel = c[c >= 1e-8]
    mu, sd = sel.mean(), np.sqrt(np.float32(sel.var()))
    expect = np.abs(c - mu) > 2.0 * sd
    assert n == expect.sum() and np.array_equal(flags.astype(bool), expect)
    flags2 = flags.copy()
    O.detect_outliers(np.ones(8), flags2, 2.0)  # nothing new, old flags stay
    assert np.array_equal(flags2, flags)
