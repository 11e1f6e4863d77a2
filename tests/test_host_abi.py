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


def _rand_pose(rng, angle):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    return rng.normal(size=3), np.concatenate([np.sin(angle / 2) * ax, [np.cos(angle / 2)]])


def test_se3_exp_log_against_matrix_exponential(pkg, cuda_lib, O):
    """mbavo_se3_exp / mbavo_se3_log (Transformation::exp / log = Sophus::SE3d::exp / log, third-party and un-vendored —
    parity unpinned): against the oracle's restatement of the published closed forms AND against scipy's matrix exponential
    of the 4 x 4 twist, from the first-order branch (theta < 1e-10) to rotations near pi; log inverts exp."""
    from scipy.linalg import expm

    from mbavo_b200 import api

    rng = np.random.default_rng(3)
    for scale in (0.0, 1e-12, 1e-9, 1e-6, 1e-3, 0.1, 1.0, 3.0):
        for _ in range(8):
            tg = np.concatenate([rng.normal(size=3), rng.normal(size=3) * scale])
            if np.linalg.norm(tg[3:]) > 3.1:
                tg[3:] *= 3.1 / np.linalg.norm(tg[3:])
            t, q = api.se3_exp(tg)
            t_o, q_o = O.se3_exp(tg)
            assert np.abs(t - t_o).max() <= 1e-14 * max(1.0, np.abs(t_o).max()) and np.abs(q - q_o).max() <= 1e-15
            M = np.zeros((4, 4))
            M[:3, :3], M[:3, 3] = O.hat(tg[3:]), tg[:3]
            E = expm(M)
            th = np.linalg.norm(tg[3:])
            # the published closed form is what it is: below theta = 1e-10 it takes V = R instead of I + hat(omega) / 2 (an
            # O(theta) difference), above it (1 - cos theta) / theta^2 cancels (an O(eps / theta) one)
            slack = 4 * min(th, 2.3e-16 / th) * max(1.0, np.abs(tg[:3]).max()) if th > 0 else 0.0
            assert np.abs(O.q_to_R(q) - E[:3, :3]).max() <= 1e-13 and np.abs(t - E[:3, 3]).max() <= 1e-12 + slack
            back = api.se3_log(t, q)
            assert np.abs(back - tg).max() <= 1e-10 * max(1.0, 1.0 / max(np.pi - th, 1e-3)) + 2 * slack
            assert np.abs(back - O.se3_log(t, q)).max() <= 1e-13


def test_spline_pose_and_transforms(pkg, cuda_lib, O, synth):
    """mbavo_spline_pose (SplineSE3::GetPose), mbavo_spline_transform_by_right / _transform_to (Spline.h:184-219),
    mbavo_predict_spline / mbavo_frame_velocity (tracker.cpp:120-161): against the numpy restatements, and through the
    properties the tracker relies on — TransformTo puts the pose at `time` on the target, a prediction with the velocity of
    the frame pair (prev, cur) maps cur's pose onto the constant-velocity continuation."""
    from conftest import reference_test_spline
    from mbavo_b200 import api

    kt, kR = reference_test_spline()
    rng = np.random.default_rng(8)
    for k in (2, 4):
        for time in (0.0, 0.26, 0.75, 1.49):
            t, q = api.spline_pose(k, 0.0, 0.5, kt, kR, time)
            t_w, q_w = synth.spline_pose(k, kt, kR, 0.0, 0.5, time)
            assert np.abs(t - t_w).max() <= 1e-13 and np.abs(q - q_w).max() <= 1e-14
        with pytest.raises(pkg.MbavoError):
            api.spline_pose(k, 0.0, 0.5, kt, kR, -0.6)   # a whole segment before the first knot: the reference asserts
        with pytest.raises(pkg.MbavoError):
            api.spline_pose(k, 0.0, 0.5, kt, kR, 0.5 * (len(kt) - k + 1) + 0.01)  # past the last segment
        # TransformTo against its restatement (every knot right-multiplied by T(time)^-1 T_target); the pose at `time` lands
        # exactly on the target only where the spline interpolates a knot (k = 2, u = 0) — the per-knot update is the reference's
        tt, tq = _rand_pose(rng, 0.7)
        for time in (0.5, 0.6):
            nt, nR = api.spline_transform_to(k, 0.0, 0.5, kt, kR, time, tt, tq)
            t0_, q0_ = synth.spline_pose(k, kt, kR, 0.0, 0.5, time)
            qi = np.array([-q0_[0], -q0_[1], -q0_[2], q0_[3]]) / (q0_ @ q0_)
            wt, wR = O.transform_by_right(kt, kR, O.q_mul(qi, tq), O.q_to_R(q0_ / np.linalg.norm(q0_)).T @ (tt - t0_))
            assert np.abs(nt - wt).max() <= 1e-12 * max(1.0, np.abs(wt).max()) and np.abs(nR - wR).max() <= 1e-14
            if k == 2 and time == 0.5:
                t, q = api.spline_pose(k, 0.0, 0.5, nt, nR, time)
                assert np.abs(t - tt).max() <= 1e-12 and min(np.abs(q - tq).max(), np.abs(q + tq).max()) <= 1e-13
    # prediction and velocity against the rotation-matrix restatement
    for _ in range(10):
        vel = np.concatenate([rng.normal(size=3), rng.normal(size=3) * 0.3])
        dtf = rng.uniform(0.01, 0.2)
        pt, pR = api.predict_spline(kt, kR, vel, dtf)
        ot, oR = O.predict_spline(kt, kR, vel, dtf)
        assert np.abs(pt - ot).max() <= 1e-12 * max(1.0, np.abs(ot).max()) and np.abs(pR - oR).max() <= 1e-15
        (t0, q0), (t1, q1) = _rand_pose(rng, 0.4), _rand_pose(rng, 0.5)
        v = api.frame_velocity(t0, q0, t1, q1, dtf)
        assert np.abs(v - O.frame_velocity(t0, q0, t1, q1, dtf)).max() <= 1e-12 * np.abs(v).max()
        # constant velocity: a one-knot "spline" at pose 1 predicted with v lands on T1 (T0^-1 T1)
        ct, cR = api.predict_spline(t1[None], q1[None], v, dtf)
        R0, R1 = O.q_to_R(q0), O.q_to_R(q1)
        want_R = R1 @ R0.T @ R1
        want_t = R1 @ (R0.T @ (t1 - t0)) + t1
        assert np.abs(O.q_to_R(cR[0]) - want_R).max() <= 1e-12 and np.abs(ct[0] - want_t).max() <= 1e-12


def test_tracker_state_and_keyframe_decision(pkg, cuda_lib, O, synth):
    """The host-only parts of the per-frame driver: mbavo_tracker_init (tracker.cpp:91-109), mbavo_tracker_new_keyframe
    (:186-199: mTKeyframe *= T(capture), the spline re-anchored so that its pose at the capture time is the identity, mTprevB2W
    reset) and mbavo_is_keyframe (:251-262)."""
    import ctypes as C

    from mbavo_b200 import api

    trk = api._Tracker()
    assert cuda_lib.mbavo_tracker_init(C.byref(trk), C.c_int(4), C.c_double(0.1), C.c_double(2.0)) != 0  # two knots: k = 2 only
    assert cuda_lib.mbavo_tracker_init(C.byref(trk), C.c_int(2), C.c_double(0.1), C.c_double(2.0)) == 0
    assert trk.num_ctrl_knots == 2 and trk.start_time == 2.0 and trk.prev_timestamp == 2.0
    assert list(trk.knots_R[:8]) == [0, 0, 0, 1, 0, 0, 0, 1] and list(trk.knots_t[:6]) == [0] * 6
    assert list(trk.keyframe_q) == [0, 0, 0, 1] and list(trk.velocity) == [0] * 6
    # put the tracker somewhere: two knots of a moving spline, a keyframe pose, then re-anchor at a time inside the segment
    rng = np.random.default_rng(2)
    kt = rng.normal(size=(2, 3))
    kR = np.array([np.concatenate([np.sin(a / 2) * ax / np.linalg.norm(ax), [np.cos(a / 2)]])
                   for a, ax in ((0.3, rng.normal(size=3)), (0.5, rng.normal(size=3)))])
    key_t, key_q = rng.normal(size=3), np.array([0.0, np.sin(0.2), 0.0, np.cos(0.2)])
    for i in range(6):
        trk.knots_t[i] = kt.reshape(-1)[i]
    for i in range(8):
        trk.knots_R[i] = kR.reshape(-1)[i]
    for i in range(3):
        trk.keyframe_t[i] = key_t[i]
    for i in range(4):
        trk.keyframe_q[i] = key_q[i]
    trk.prev_t[0] = 5.0
    cap = 2.04
    t_c, q_c = synth.spline_pose(2, kt, kR, 2.0, 0.1, cap)
    assert cuda_lib.mbavo_tracker_new_keyframe(C.byref(trk), C.c_double(cap)) == 0
    want_t = O.q_to_R(key_q) @ t_c + key_t
    want_q = O.q_mul(key_q, q_c)
    assert np.abs(np.array(trk.keyframe_t) - want_t).max() <= 1e-14 and np.abs(np.array(trk.keyframe_q) - want_q).max() <= 1e-15
    nt, nR = np.array(trk.knots_t[:6]).reshape(2, 3), np.array(trk.knots_R[:8]).reshape(2, 4)
    t_n, q_n = synth.spline_pose(2, nt, nR, 2.0, 0.1, cap)
    # exactly the identity in rotation; in translation up to the reference's per-knot update (Spline.h:196-200)
    assert np.abs(q_n - [0, 0, 0, 1]).max() <= 1e-14 and np.abs(t_n).max() <= 0.05 * np.abs(kt).max()
    assert list(trk.prev_t) == [0, 0, 0] and list(trk.prev_q) == [0, 0, 0, 1]
    # the relative motion along the spline is what it was, seen from the new body frame (conjugated by the re-anchoring transform)
    before = O.frame_velocity(*synth.spline_pose(2, kt, kR, 2.0, 0.1, 2.0), *synth.spline_pose(2, kt, kR, 2.0, 0.1, 2.1), 1.0)
    after = O.frame_velocity(*synth.spline_pose(2, nt, nR, 2.0, 0.1, 2.0), *synth.spline_pose(2, nt, nR, 2.0, 0.1, 2.1), 1.0)
    assert abs(np.linalg.norm(before[3:]) - np.linalg.norm(after[3:])) <= 1e-12
    # keyframe decision
    f = cuda_lib.mbavo_is_keyframe
    f.argtypes = [C.c_double] * 5
    assert f(12.0, 3.0, 10.0, 30.0, 5.0) == 1      # far enough and sharp enough
    assert f(12.0, 6.0, 10.0, 30.0, 5.0) == 0      # far enough but too blurred
    assert f(31.0, 6.0, 10.0, 30.0, 5.0) == 1      # too far, whatever the blur
    assert f(9.0, 1.0, 10.0, 30.0, 5.0) == 0


def test_header_is_plain_c_and_example_links(pkg, cuda_lib, tmp_path):
    """include/mbavo.h is a C header (no C++ in the signatures): the plain-C example compiles with gcc -std=c99 -Wall -Werror and
    links against the product library (running it needs a GPU: tests/test_gpu_tracking.py)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "mba-vo_b200", "lib")
    exe = os.path.join(str(tmp_path), "track_sequence")
    r = subprocess.run(["/usr/bin/gcc", "-O2", "-std=c99", "-Wall", "-Werror", os.path.join(root, "examples", "track_sequence.c"),
                        "-I" + os.path.join(root, "include"), "-L" + libdir, "-lmbavo_b200", "-lm", "-Wl,-rpath," + libdir,
                        "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_pose_only_series_host(tmp_path):
    """The pose-only path of the persistent sweep (csrc/pose_device.cuh: so3_log_only / so3_exp_only / spline_pose_only — power
    series in the squared norm instead of sqrt / atan / sincos — and sample_u) compiled as HOST code: against long-double closed
    forms of Exp / Log over every branch (tiny, series, closed form), against the closed-form path spline_pose keeps (the
    restatement of SplineFunctor.h:155-365) for linear and cubic segments, and against the plain expression of the sample
    position (compute_virtual_camera_poses.cu:33).  All at the fp64 rounding level."""
    import os
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(str(tmp_path), "pose_series")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-x", "cu", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", exe, os.path.join(root, "tests", "cpp", "pose_series_main.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    err = {k: float(v) for k, v in (line.split() for line in out.stdout.strip().splitlines())}
    assert err["exp_only_vs_long_double"] <= 5e-16 and err["log_only_vs_long_double"] <= 1e-15 and err["log_exp_round_trip"] <= 1e-15, err
    assert err["pose_only_vs_closed_form_t"] == 0.0 and err["pose_only_vs_closed_form_q"] <= 5e-16, err
    assert err["sample_u_vs_plain"] <= 1e-12, err
