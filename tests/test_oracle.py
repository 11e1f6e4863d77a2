"""CPU tests of the oracle: pinned against oracle/_ref (the reference's own header arithmetic) where it is built,
against the committed golden vectors everywhere, and against the eight self-consistency checks of the reference's
test executable (test/test_blur_aware_tracker_modules.cpp) restated as ASSERTING tests."""
import numpy as np
import pytest

from conftest import reference_test_spline
from helpers import first_step, golden, max_rel, problem_from_golden, rel


# ----------------------------------------------------------------------------------------------------------------
# golden vectors (generated from oracle/_ref by tests/golden/make_golden.py)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 4])
def test_golden_virtual_poses(orc, k):
    z = golden(f"virtual_poses_k{k}.npz")
    poses, seg, Jt, JR = orc.virtual_poses(int(z["N"]), z["cap"], z["exp"], k, float(z["t0"]), float(z["dt"]), z["knots_t"],
                                           z["knots_R"])
    assert np.array_equal(seg, z["seg"])
    assert np.abs(poses - z["poses"]).max() <= 1e-13
    assert np.abs(Jt - z["Jt"]).max() <= 1e-13
    assert np.abs(JR - z["JR"]).max() <= 1e-12  # SURVEY §8d gate: poses / J <= 1e-10


def test_golden_pixel_intensity(orc, synth):
    z = golden("pixel_intensity.npz")
    I = synth.ramp_image(int(z["H"]), int(z["W"]))
    g = synth.image_gradient(I)
    n_valid = 0
    for row in z["rows"]:
        pose, (D, X, Y, ok_ref, I_ref), J_ref = row[:7], row[7:12], row[12:]
        ok, inten, J = orc.pixel_intensity(I, g, pose, D, float(z["fx"]), float(z["fy"]), float(z["cx"]), float(z["cy"]), X, Y)
        assert ok == bool(ok_ref)
        if ok:
            n_valid += 1
            assert abs(inten - I_ref) <= 1e-4          # fp32 bilinear: FMA contraction may differ between builds
            assert np.abs(J - J_ref).max() <= 1e-4 * max(1.0, np.abs(J_ref).max())
    assert n_valid >= 32


@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_golden_evaluate(orc, O, synth, tag):
    z = golden(f"evaluate_{tag}.npz")
    prob = problem_from_golden(z, synth)
    c, H, g, pc = orc.evaluate(prob, 0)
    assert abs(c - float(z["cost"])) <= 1e-7 * abs(float(z["cost"]))
    assert max_rel(H, z["Hessian"]) <= 1e-6 and max_rel(g, z["gradient"]) <= 1e-6
    assert np.abs(pc - z["patch_costs"]).max() <= 1e-6
    assert rel(first_step(O, H, g), first_step(O, z["Hessian"].copy(), z["gradient"])) <= 1e-5
    c2 = orc.evaluate(prob, 0, with_hessian=False)[0]
    assert abs(c2 - float(z["cost_only"])) <= 1e-7 * abs(float(z["cost_only"]))
    # outlier flags: skipped in the sums, counted out of the normaliser (…cost.cu:267, spline_update_step.cpp:116)
    flags = z["flags"]
    cf, Hf, gf, pcf = orc.evaluate(prob, 0, flags=flags, num_bad=int(flags.sum()))
    assert abs(cf - float(z["cost_flagged"])) <= 1e-7 * abs(float(z["cost_flagged"]))
    assert max_rel(Hf, z["H_flagged"]) <= 1e-6 and max_rel(gf, z["g_flagged"]) <= 1e-6
    assert np.abs(pcf - z["patch_costs_flagged"]).max() <= 1e-6


def test_golden_blurred(orc):
    z = golden("blurred.npz")
    out = orc.synthesize_blurred(np.ascontiguousarray(z["ref_I"]), float(z["D"]), float(z["fx"]), float(z["fy"]),
                                 float(z["cx"]), float(z["cy"]), 2, 0.0, 1.0, z["knots_t"], z["knots_R"], float(z["cap"]),
                                 float(z["exp"]), int(z["n"]))
    diff = np.abs(out.astype(int) - z["out"].astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3  # a float-rounding tie may move a pixel by one grey level


def test_golden_pyramid_and_gradients(orc, O, synth):
    """The 2 x 2 box pyramid and the central-difference gradients as the reference's own loops produced them
    (tests/golden/make_golden.py, section 6): the numpy restatements (package synth, oracle) and the oracle's C
    restatement reproduce every level bit for bit — these are the loops the GPU pyramid path is compared with."""
    z = golden("pyramid.npz")
    cur = np.ascontiguousarray(z["I0"])
    levels = O.pyramid_numpy(cur, 3)
    for l in range(3):
        if l > 0:
            assert np.array_equal(np.asarray(orc.pyramid_down(cur)), z[f"I{l}"])
            cur = synth.pyramid_down(cur)
        assert np.array_equal(cur, z[f"I{l}"]) and np.array_equal(levels[l], z[f"I{l}"])
        g = z[f"g{l}"]
        assert np.array_equal(synth.image_gradient(cur).reshape(g.shape), g)
        assert np.array_equal(np.asarray(orc.image_gradient(cur)).reshape(g.shape), g)


def _selection_cases():
    z = golden("point_selection.npz")
    for name in ("tex", "ramp"):
        cnt = z[name + "_count"]
        offs = np.concatenate([[0], np.cumsum(cnt)])
        want = [(z[name + "_xy"][offs[l]:offs[l + 1]], z[name + "_z"][offs[l]:offs[l + 1]]) for l in range(len(cnt))]
        yield name, np.ascontiguousarray(z[name + "_I"]), np.ascontiguousarray(z[name + "_depth"]), float(z[name + "_thr"]), int(z[name + "_cell"]), want


def test_golden_point_selection(O):
    """numpy restatement of FeatureDetectorSemiDense::detect + gridSelection + the depth look-up against the points the
    reference's own detector sources selected (tests/golden/make_golden.py, section 5): bit-exact, same order."""
    for name, I, depth, thr, cell, want in _selection_cases():
        got = O.select_points(I, len(want), thr, cell, cell, depth)
        assert sum(len(zz) for _, zz in want) > 100
        for (xy, zz), (xy_w, zz_w) in zip(got, want):
            assert np.array_equal(xy, xy_w) and np.array_equal(zz, zz_w), name


# ----------------------------------------------------------------------------------------------------------------
# directly against oracle/_ref (skipped where it is not built)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kw", [("tiny", {}), ("C5cubic", dict(W=160, H=120, P0=200, N=8, n_knots=4))])
def test_oracle_matches_reference_arithmetic(orc, ref, O, synth, name, kw):
    prob = synth.make_config(name, **kw)
    for level in range(len(prob.levels)):
        c1, H1, g1, p1 = orc.evaluate(prob, level)
        c2, H2, g2, p2 = ref.evaluate(prob, level)
        assert abs(c1 - c2) <= 1e-12 * abs(c2)
        assert max_rel(H1, H2) <= 1e-10 and max_rel(g1, g2) <= 1e-10 and np.abs(p1 - p2).max() <= 1e-12
        assert rel(first_step(O, H1, g1), first_step(O, H2, g2)) <= 1e-8


def test_oracle_matches_reference_poses_and_centres(orc, ref):
    kt, kR = reference_test_spline()
    cap, exp = 0.25 + 0.5 * np.arange(4), np.full(4, 0.1)
    for k in (2, 4):
        a = orc.virtual_poses(32, cap, exp, k, 0.0, 0.5, kt, kR)
        b = ref.virtual_poses(32, cap, exp, k, 0.0, 0.5, kt, kR)
        assert np.array_equal(a[1], b[1])
        for x, y in ((a[0], b[0]), (a[2], b[2]), (a[3], b[3])):
            assert np.abs(x - y).max() <= 1e-13
        rng = np.random.default_rng(3)
        xy = np.stack([rng.uniform(20, 620, 145), rng.uniform(20, 460, 145)], axis=1)  # test_compute_local_patches :344-500
        z = rng.uniform(20, 45, 145)
        ca = orc.local_patches(32, a[0], xy, z, 320, 320, 320, 240)
        cb = ref.local_patches(32, b[0], xy, z, 320, 320, 320, 240)
        assert np.abs(ca - cb).max() <= 1e-9  # SURVEY §8d gate: patch centres <= 1e-9 px


# ----------------------------------------------------------------------------------------------------------------
# the reference's own module checks, restated with assertions
# ----------------------------------------------------------------------------------------------------------------
def test_ref_check_pixel_intensity_round_trip(orc, synth):
    """test_compute_pixel_intensity (:83-181): warp of the un-projected reference pixel lands back on it, so the warped
    intensity equals the bilinear sample at ref_xy = (20.5, 20.5); analytic vs numeric 1x7 Jacobian (eps 1e-3: with
    the test's 1e-6 the fp32 bilinear weights make the numeric Jacobian pure noise)."""
    I = synth.ramp_image(480, 640)
    g = synth.image_gradient(I)
    fx = fy = 320.0
    cx, cy = 320.0, 240.0
    rng = np.random.default_rng(0)
    for _ in range(8):
        D = rng.uniform(5, 10)
        q = rng.normal(size=4) * np.array([0.1, 0.1, 0.1, 1.0])
        q /= np.linalg.norm(q)
        t = rng.uniform(-0.5, 0.5, 3)
        R = synth.q_to_R(q)
        P_ref = D * np.array([(20.5 - cx) / fx, (20.5 - cy) / fy, 1.0])
        P_cur = R.T @ (P_ref - t)
        X, Y = fx * P_cur[0] / P_cur[2] + cx, fy * P_cur[1] / P_cur[2] + cy
        ok, inten, J = orc.pixel_intensity(I, g, np.concatenate([t, q]), D, fx, fy, cx, cy, X, Y)
        assert ok
        expect = 0.25 * (float(I[20, 20]) + float(I[20, 21]) + float(I[21, 20]) + float(I[21, 21]))
        assert abs(inten - expect) <= 1e-4
        eps = 1e-3
        for a in range(7):
            d = np.zeros(7)
            d[a] = eps
            ip = orc.pixel_intensity(I, g, np.concatenate([t, q]) + d, D, fx, fy, cx, cy, X, Y, want_J=False)[1]
            im = orc.pixel_intensity(I, g, np.concatenate([t, q]) - d, D, fx, fy, cx, cy, X, Y, want_J=False)[1]
            assert abs((ip - im) / (2 * eps) - J[a]) <= 2e-2 * max(1.0, abs(J[a]))


@pytest.mark.parametrize("k", [2, 4])
def test_ref_check_virtual_poses(orc, O, synth, k):
    """test_compute_virtual_camera_poses (:183-342): pose == independent spline evaluation, J_t / J_R == finite
    differences under t_j += e, R_j <- R_j Exp(e)."""
    kt, kR = reference_test_spline()
    cap, exp = 0.25 + 0.5 * np.arange(4), np.full(4, 0.1)
    N, t0, dt = 32, 0.0, 0.5
    poses, seg, Jt, JR = orc.virtual_poses(N, cap, exp, k, t0, dt, kt, kR)
    for f, i in ((0, 0), (2, 17), (3, 31)):
        ts = cap[f] - 0.5 * exp[f] + i * exp[f] / (N - 1 + 1e-8)
        tt, q = synth.spline_pose(k, kt, kR, t0, dt, ts)
        assert np.abs(poses[f, i, :3] - tt).max() <= 1e-12 and np.abs(poses[f, i, 3:] - q).max() <= 1e-12
        idx, eps = seg[f, i], 1e-6
        for j in range(k):
            for c in range(3):
                d = np.zeros(6 * len(kt))
                d[3 * (idx + j) + c] = eps
                tp, _ = synth.spline_pose(k, *O.plus(kt, kR, d), t0, dt, ts)
                tm, _ = synth.spline_pose(k, *O.plus(kt, kR, -d), t0, dt, ts)
                assert np.abs((tp - tm) / (2 * eps) - Jt[f, i, :, 3 * j + c]).max() <= 1e-8
                d = np.zeros(6 * len(kt))
                d[3 * len(kt) + 3 * (idx + j) + c] = eps
                _, qp = synth.spline_pose(k, *O.plus(kt, kR, d), t0, dt, ts)
                _, qm = synth.spline_pose(k, *O.plus(kt, kR, -d), t0, dt, ts)
                assert np.abs((qp - qm) / (2 * eps) - JR[f, i, :, 3 * j + c]).max() <= 1e-7


def test_ref_check_local_patches(orc, synth):
    """test_compute_local_patches (:344-500): centre == projection of the back-projected point at pose N/2."""
    kt, kR = reference_test_spline()
    cap, exp = 0.25 + 0.5 * np.arange(4), np.full(4, 0.1)
    poses, _, _, _ = orc.virtual_poses(32, cap, exp, 4, 0.0, 0.5, kt, kR, jac=False)
    rng = np.random.default_rng(1)
    xy = np.stack([rng.uniform(20, 620, 145), rng.uniform(20, 460, 145)], axis=1)
    z = rng.uniform(20, 45, 145)
    cen = orc.local_patches(32, poses, xy, z, 320, 320, 320, 240)
    for f in range(4):
        t, q = poses[f, 16, :3], poses[f, 16, 3:]
        Pr = np.stack([z * (xy[:, 0] - 320) / 320, z * (xy[:, 1] - 240) / 320, z], axis=1)
        Pc = (Pr - t) @ synth.q_to_R(q)
        expect = np.stack([Pc[:, 0] / Pc[:, 2] * 320 + 320, Pc[:, 1] / Pc[:, 2] * 320 + 240], axis=1)
        assert np.abs(cen[f] - expect).max() <= 1e-9


@pytest.mark.parametrize("name,kw", [("k2", dict(n_knots=2, k=2)), ("k2_3knots", dict(n_knots=3, k=2)),
                                     ("k4", dict(n_knots=4, k=4)), ("k4_2seg", dict(n_knots=5, k=4))])
def test_ref_check_pixel_residual_and_jacobian(orc, O, synth, name, kw):
    """test_compute_pixel_jacobian_residual (:502-895): residual == mean of warped intensities - live pixel, analytic
    Jacobian row vs numeric (patch centres held fixed), including exposure windows that straddle a knot."""
    # the reference test's ramp image: its gradient image IS the derivative of the bilinear interpolant (away from the
    # 254 -> 0 wrap), so a single pixel's numeric Jacobian is meaningful
    prob = synth.make_problem(name, W=160, H=120, levels=1, P0=40, N=8, seed=11, margin=12, image="ramp", **kw)
    lv = prob.levels[0]
    r, J, kmin, NK = orc.pixel_residuals(prob, 0)
    assert NK == prob.n_knots and kmin == 0
    poses, seg, _, _ = orc.virtual_poses(lv.N, prob.cap, prob.exp, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R)
    cen = orc.local_patches(lv.N, poses, lv.xy, lv.z, lv.fx, lv.fy, lv.cx, lv.cy)
    p, j = int(np.argmax(lv.xy.sum(1) < 180)), 1
    X, Y = int(cen[0, p, 0] + lv.pattern[j, 0]), int(cen[0, p, 1] + lv.pattern[j, 1])

    def residual(kt, kR):
        ps = orc.virtual_poses(lv.N, prob.cap, prob.exp, prob.k, prob.t0, prob.dt, kt, kR, jac=False)[0]
        s = sum(orc.pixel_intensity(lv.ref_I, lv.ref_dIxy, ps[0, i], lv.z[p], lv.fx, lv.fy, lv.cx, lv.cy, X, Y, False)[1]
                for i in range(lv.N))
        return s / lv.N - float(lv.cur_I[0][Y, X])

    assert abs(residual(prob.knots_t, prob.knots_R) - r[0, p, j]) <= 1e-9
    n = prob.n_knots
    for c in range(6 * n):
        eps = 1e-3 if c < 3 * n else 1e-4
        d = np.zeros(6 * n)
        d[c] = eps
        num = (residual(*O.plus(prob.knots_t, prob.knots_R, d)) - residual(*O.plus(prob.knots_t, prob.knots_R, -d))) / (2 * eps)
        assert abs(num - J[0, p, j, c]) <= 3e-2 * max(1.0, np.abs(J[0, p, j]).max()), (c, num, J[0, p, j, c])


def test_ref_check_patch_frame_and_merge_layout(orc, O, synth):
    """test_compute_patch_cost_gradient_hessian / _frame_ / test_merge (:897-1181): H = sum w J^T J, g = sum w r J, cost =
    sum rho, all / num_residuals, in the unknown ordering [t-block | w-block], from the per-pixel rows."""
    prob = synth.make_problem("layout", W=160, H=120, levels=1, P0=50, N=8, n_knots=3, k=2, seed=5, margin=12, huber_a=0.5)
    r, J, kmin, NK = orc.pixel_residuals(prob, 0)
    c, H, g, pc = orc.evaluate(prob, 0)
    a = prob.huber_a
    x = 0.5 * r * r
    w = np.where(x > a * a, a / (np.sqrt(x) + 1e-8), 1.0)
    rho = np.where(x > a * a, 2 * a * np.sqrt(x) - a * a, x)
    nres = r.size
    assert (x > a * a).any() and (x <= a * a).any()
    assert abs(rho.sum() / nres - c) <= 1e-6 * c                      # float sqrt in the reference: 1e-7 relative
    Jf, wf, rf = J.reshape(-1, J.shape[-1]), w.reshape(-1), r.reshape(-1)
    assert max_rel((Jf * (wf * rf)[:, None]).sum(0) / nres, g) <= 1e-6
    assert max_rel((Jf.T * wf) @ Jf / nres, H) <= 1e-6
    assert np.abs(rho.sum(-1) / nres - pc).max() <= 1e-6 * pc.max()
    assert np.abs(H - H.T).max() == 0.0


def test_ref_check_merge_overlapping_frames(orc, synth):
    """test_merge_hessian_gradient_cost (:1060-1181): several frames whose segments overlap accumulate into the same
    global blocks; index 3*knot for t, 3*(n+knot) for w.  Evaluating F frames at once == sum of single-frame runs."""
    prob = synth.make_problem("frames", W=160, H=120, levels=1, P0=30, N=4, n_knots=3, k=2, seed=8, margin=12)
    lv = prob.levels[0]
    prob.dt = 1.0
    prob.cap, prob.exp = np.array([0.5, 1.45]), np.array([0.8, 0.8])  # frame 0 in segment 0, frame 1 in segment 1
    lv.cur_I = [lv.cur_I[0], np.ascontiguousarray(lv.cur_I[0][::-1])]
    c, H, g, _ = orc.evaluate(prob, 0)
    tot_c, tot_H, tot_g = 0.0, 0.0, 0.0
    for f in range(2):
        single = synth.Problem(**{**prob.__dict__, "cap": prob.cap[f:f + 1], "exp": prob.exp[f:f + 1]})
        single.levels = [synth.Level(**{**lv.__dict__, "cur_I": [lv.cur_I[f]]})]
        cf, Hf, gf, _ = orc.evaluate(single, 0)
        tot_c, tot_H, tot_g = tot_c + cf / 2, tot_H + Hf / 2, tot_g + gf / 2  # normaliser counts F frames
    assert abs(c - tot_c) <= 1e-12 * c and np.abs(H - tot_H).max() <= 1e-12 * np.abs(H).max()
    assert np.abs(g - tot_g).max() <= 1e-12 * np.abs(g).max()
    n = 3
    assert np.abs(H[0:3, 6:9]).max() == 0.0          # knot 0 and knot 2 never share a frame (t-t block)
    assert np.abs(H[3:6, 3:6]).max() > 0.0 and np.abs(H[3 * n + 3:3 * n + 6, 3:6]).max() > 0.0


def test_ref_check_solve_normal_equation(O):
    """test_solve_normal_equation (:1183-1208) for both branches: x = -A^-1 b on a random SPD 48x48."""
    rng = np.random.default_rng(2)
    M = rng.normal(size=(48, 48))
    A = M @ M.T + 48 * np.eye(48)
    b = rng.normal(size=48)
    for solver in ("SVD_JACOBI", "LDLT"):
        x = O.solve_normal_equation(A, b, solver)
        assert np.abs(A @ x + b).max() <= 1e-8


def test_cost_gradient_consistency_multi_segment(orc, O, synth):
    """g is the gradient of the cost when the patch centres are held fixed — checks the per-sample segment scatter that
    the reference lacks by central differences of the cost."""
    for kw in (dict(n_knots=3, k=2), dict(n_knots=5, k=2), dict(n_knots=6, k=4)):
        prob = synth.make_problem("fd", W=320, H=240, levels=1, P0=1500, N=16, seed=21, **kw)
        prob.huber_a = 1e6
        cen = np.zeros((1, prob.levels[0].P, 2))
        c, H, g, _ = orc.evaluate(prob, 0, centres_out=cen)
        n = prob.n_knots
        gfd = np.zeros(6 * n)
        for j in range(6 * n):
            eps = 1e-3 if j < 3 * n else 1e-4
            d = np.zeros(6 * n)
            d[j] = eps
            cp = orc.evaluate(prob, 0, *O.plus(prob.knots_t, prob.knots_R, d), with_hessian=False, centres_in=cen)[0]
            cm = orc.evaluate(prob, 0, *O.plus(prob.knots_t, prob.knots_R, -d), with_hessian=False, centres_in=cen)[0]
            gfd[j] = (cp - cm) / (2 * eps)
        assert np.abs(g - gfd).max() <= 5e-2 * np.abs(g).max(), kw


def test_input_format_helpers(orc, synth):
    """Gradient.h:17-75 and ImagePyramid.h:59-99: the oracle's C restatement == the numpy generator used for inputs."""
    I = synth.make_texture(61, 83, 3)
    assert np.array_equal(orc.image_gradient(I), synth.image_gradient(I))
    assert np.array_equal(orc.pyramid_down(I), synth.pyramid_down(I))
    g = synth.image_gradient(I)
    assert np.all(g[0] == 0) and np.all(g[-1] == 0) and np.all(g[:, 0] == 0) and np.all(g[:, -1] == 0)


@pytest.mark.parametrize("thr,cell", [(25.0, (30, 30)), (5.0, (30, 30)), (0.5, (17, 23)), (60.0, (8, 8))])
def test_oracle_matches_reference_point_selection(O, synth, thr, cell):
    """Directly against the reference's detector (oracle/_ref/libmbavo_refselect.so): textured, ramp (ties), noise and flat
    VGA images over 4 levels, the tracker's parameters (25, 30 x 30) and others; also the magnitude image itself."""
    if not O.RefSelect.available():
        pytest.skip("oracle/_ref/libmbavo_refselect.so not built (no /root/reference here)")
    rs = O.RefSelect()
    rng = np.random.default_rng(5)
    H, W = 480, 640
    images = [synth.make_config("C1").levels[0].ref_I, synth.ramp_image(H, W), rng.integers(0, 256, (H, W)).astype(np.uint8),
              np.full((H, W), 77, np.uint8)]
    for I in images:
        depth = rng.uniform(0.0, 10.0, (H, W)).astype(np.float32)
        depth[rng.random((H, W)) < 0.1] = 0.0
        want, mag = rs.select_points(I, 4, thr, cell[0], cell[1], depth, want_mag=True)
        got = O.select_points(I, 4, thr, cell[0], cell[1], depth)
        assert np.array_equal(mag, O.gradient_magnitude(I))
        for (xy, zz), (xy_w, zz_w) in zip(got, want):
            assert np.array_equal(xy, xy_w) and np.array_equal(zz, zz_w)


def test_oracle_matches_reference_pyramid(orc, O, synth):
    """Directly against the reference's ImagePyramid<T>::computePyramid and compute_image_gradients (oracle/_ref)."""
    if not O.RefSelect.available():
        pytest.skip("oracle/_ref/libmbavo_refselect.so not built (no /root/reference here)")
    rs = O.RefSelect()
    rng = np.random.default_rng(1)
    for H, W, L in ((123, 161, 4), (480, 640, 4), (47, 33, 3)):
        I = rng.integers(0, 256, (H, W)).astype(np.uint8)
        cur = I
        for l, (im, g) in enumerate(rs.pyramid(I, L)):
            if l > 0:
                assert np.array_equal(np.asarray(orc.pyramid_down(cur)), im)
                cur = synth.pyramid_down(cur)
            assert np.array_equal(cur, im)
            assert np.array_equal(synth.image_gradient(cur).reshape(g.shape), g)
            assert np.array_equal(np.asarray(orc.image_gradient(cur)).reshape(g.shape), g)


def test_shapes_image_restatement_matches_opencv(synth):
    """synth.shapes_image (the reference's synthetic scene, generate_synthetic_data.cpp:11-125) against the image OpenCV's own
    cv::fillPoly drew from the same vertices (tests/golden/shapes.npz, written by tests/golden/make_golden.py through cv2)."""
    z = golden("shapes.npz")
    assert np.array_equal(synth.shapes_image(480, 640), z["image"])
    assert set(np.unique(z["image"])) == {0, 255} and 60000 < int((z["image"] == 255).sum()) < 70000
