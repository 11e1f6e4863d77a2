"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs and against the committed golden vectors.

Gates (BASELINE.md §3.4 — the north star's floating-point tolerances):
    first LM step   ||delta - delta_ref|| / ||delta_ref|| <= 1e-4     (radius 1e4)
    cost            |c - c_ref| / |c_ref|                 <= 1e-5
    H, g            max |x - x_ref| / max |x_ref|         <= 1e-5 (element-wise gate of §8d)
The 1e-4 gate on the step holds for every BASELINE config; for deliberately degenerate cases (one or three points, the
60-point cubic golden case with cond(H) = 2.5e10) helpers.delta_gate widens it to the reproducibility of the reference's
own fp32 sampling, measured by perturbing the oracle's H, g by one fp32 ulp.
"""
import numpy as np
import pytest

from helpers import delta_gate, first_step, golden, max_rel, problem_from_golden, rel
from helpers import run_ranks as _run_ranks

pytestmark = pytest.mark.gpu

COST_TOL = 1e-5
DELTA_TOL = 1e-4


@pytest.fixture(scope="module")
def api(pkg):
    from mbavo_b200 import api as a

    return a


def gpu_eval(pkg, api, prob, level=0, knots=None, with_hessian=True, flags=None, num_bad=0, ctx=None):
    own = ctx is None
    if own:
        ctx = pkg.Context(api.limits_for(prob))
        api.upload_problem(ctx, prob)
    try:
        if flags is not None:
            ctx.set_outliers(level, flags, num_bad)
        kt, kR = (prob.knots_t, prob.knots_R) if knots is None else knots
        c, H, g = ctx.evaluate(level, prob.k, prob.t0, prob.dt, kt, kR, prob.huber_a, with_hessian)
        pc = ctx.patch_costs(level, prob.F, prob.levels[level].P)
        return c, H, g, pc
    finally:
        if own:
            ctx.close()


def check_parity(O, got, want, cost_tol=COST_TOL, delta_tol=DELTA_TOL, strict=False):
    c, H, g, pc = got
    c_ref, H_ref, g_ref, pc_ref = want
    assert abs(c - c_ref) <= cost_tol * abs(c_ref), ("cost", c, c_ref)
    assert np.abs(pc - pc_ref).max() <= 1e-4 * max(pc_ref.max(), 1e-12), "patch costs"
    if H is not None:
        assert np.array_equal(H, H.T)
        # element-wise gate of SURVEY §8d: 1e-5 of the largest element
        assert max_rel(H, H_ref) <= 1e-5 and max_rel(g, g_ref) <= 1e-5, (max_rel(H, H_ref), max_rel(g, g_ref))
        d, d_ref = first_step(O, H, g), first_step(O, H_ref, g_ref)
        # 1e-4, relaxed only where the reference's own fp32 sampling makes ITS step less reproducible than that
        gate = delta_tol if strict else delta_gate(O, H_ref, g_ref, base=delta_tol)
        assert rel(d, d_ref) <= gate, ("delta", rel(d, d_ref), gate)


# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_golden_reference_vectors(pkg, api, O, synth, tag):
    """Against outputs of the reference's own arithmetic (tests/golden, generated from oracle/_ref)."""
    z = golden(f"evaluate_{tag}.npz")
    prob = problem_from_golden(z, synth)
    check_parity(O, gpu_eval(pkg, api, prob), (float(z["cost"]), z["Hessian"], z["gradient"], z["patch_costs"]))
    c2 = gpu_eval(pkg, api, prob, with_hessian=False)[0]
    assert abs(c2 - float(z["cost_only"])) <= COST_TOL * float(z["cost_only"])
    flags = z["flags"]
    check_parity(O, gpu_eval(pkg, api, prob, flags=flags, num_bad=int(flags.sum())),
                 (float(z["cost_flagged"]), z["H_flagged"], z["g_flagged"], z["patch_costs_flagged"]))


@pytest.mark.parametrize("name,levels", [("tiny", None), ("C1", None), ("C2", None), ("C3", None), ("C5", None),
                                         ("C5cubic", None)])
def test_baseline_configs_against_oracle(pkg, api, O, orc, synth, name, levels):
    """BASELINE.json configs, every pyramid level (the oracle runs on all host cores): Hessian pass and cost-only pass."""
    import os

    orc.set_num_threads(len(os.sched_getaffinity(0)))
    prob = synth.make_config(name)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        for level in (levels if levels is not None else range(len(prob.levels))):
            want = orc.evaluate(prob, level)
            check_parity(O, gpu_eval(pkg, api, prob, level, ctx=ctx), want, strict=True)  # the north-star gate, unrelaxed
            c2 = gpu_eval(pkg, api, prob, level, with_hessian=False, ctx=ctx)
            assert abs(c2[0] - want[0]) <= COST_TOL * want[0]
            assert np.abs(c2[3] - want[3]).max() <= 1e-4 * want[3].max()


@pytest.mark.parametrize("P,S,N", [(1, 8, 4), (3, 8, 5), (301, 8, 7), (257, 1, 8), (130, 5, 3), (77, 12, 16), (33, 40, 9),
                                   (64, 33, 1), (500, 8, 64), (129, 128, 2)])
def test_ragged_shapes(pkg, api, O, orc, synth, P, S, N):
    """Point counts that do not fill a warp batch, patch sizes that are not 8 / not a power of two / larger than a warp,
    sample counts that are not powers of two (the reference requires powers of two: reduction.h:13-55)."""
    rng = np.random.default_rng(P * 1000 + S * 10 + N)
    pattern = None
    if S != 8:
        pattern = np.stack([rng.integers(-3, 4, S), rng.integers(-3, 4, S)], axis=1).astype(np.int32)
    prob = synth.make_problem("ragged", W=192, H=144, levels=1, P0=P, N=N, n_knots=2, k=2, seed=P + S + N, pattern=pattern,
                              margin=14)
    check_parity(O, gpu_eval(pkg, api, prob), orc.evaluate(prob, 0))
    c2 = gpu_eval(pkg, api, prob, with_hessian=False)[0]
    assert abs(c2 - orc.evaluate(prob, 0, with_hessian=False)[0]) <= COST_TOL * c2


def test_image_borders_and_invalid_samples(pkg, api, O, orc, synth):
    """Points right up to the image border: pixels outside the live image give r = 0, J = 0; samples that leave the
    keyframe contribute 0 with the divisor kept at N (…cost.cu:74-78, 107-121)."""
    prob = synth.make_problem("border", W=160, H=120, levels=1, P0=4000, N=8, n_knots=2, k=2, seed=77, margin=0, motion_scale=2.0)
    lv = prob.levels[0]
    r, _, _, _ = orc.pixel_residuals(prob, 0, with_jacobian=False)
    assert (r == 0).sum() > 50, "the case must actually contain invalid pixels"
    got, want = gpu_eval(pkg, api, prob), orc.evaluate(prob, 0)
    # a sample within fp32 rounding of the validity boundary may flip: allow a few patches to differ, nothing else
    diff = np.abs(got[3] - want[3])
    assert (diff > 1e-4 * want[3].max()).mean() < 2e-3
    assert abs(got[0] - want[0]) <= 1e-4 * want[0]
    assert rel(first_step(O, got[1], got[2]), first_step(O, want[1], want[2])) <= 1e-3
    # last row / column taps (weight 0) must not read out of bounds: a point whose warp lands on x = W-1 exactly
    prob2 = synth.make_problem("edge", W=64, H=48, levels=1, P0=1, N=2, n_knots=2, k=2, seed=1, margin=4)
    prob2.knots_t[:] = 0
    prob2.knots_R[:] = [0, 0, 0, 1]
    prob2.levels[0].xy[0] = [62.9999, 46.9999]
    prob2.levels[0].pattern = np.array([[0, 0], [1, 1], [0, 1], [1, 0]], dtype=np.int32)
    check_parity(O, gpu_eval(pkg, api, prob2), orc.evaluate(prob2, 0), cost_tol=1e-5, delta_tol=1.0)


def test_reference_shapes_scene(pkg, api, O, orc, synth):
    """The reference's own synthetic scene (synthesize_img_with_rand_shapes, generate_synthetic_data.cpp:11-125: white rectangles
    and triangles on black, i.e. zero gradient almost everywhere and 127.5-per-pixel steps on the edges) blurred along a spline
    as synthesize_motion_blurred_img does: evaluation parity on semi-dense points picked on its edges by the GPU selection, which
    must be the oracle's selection."""
    img = synth.shapes_image(480, 640)
    assert np.array_equal(img, golden("shapes.npz")["image"])
    prob = synth.make_problem("shapes", W=640, H=480, levels=2, P0=8, N=16, n_knots=2, k=2, seed=3, image="shapes", depth_mode="plane")
    depth = np.full((480, 640), 7.5, dtype=np.float32)
    lv0 = prob.levels[0]
    with pkg.Context(api.Limits(max_num_keypoints=4096, max_num_virtual_poses_per_frame=16, max_patch_size=8, max_num_ctrl_knots=2)) as ctx:
        ctx.set_frame_times(prob.cap, prob.exp)
        ctx.set_keyframe_pyramid(2, lv0.ref_I)
        ctx.set_live_pyramid(2, lv0.cur_I)
        counts = ctx.select_points(2, depth, lv0.fx, lv0.fy, lv0.cx, lv0.cy, lv0.pattern, lv0.N, 25.0, 30, 30)
        want_pts = O.select_points(lv0.ref_I, 2, 25.0, 30, 30, depth)
        assert counts == [len(z) for _, z in want_pts] and counts[0] > 50
        for level in range(2):
            xy, z = ctx.get_points(level)
            assert np.array_equal(xy, want_pts[level][0]) and np.array_equal(z, want_pts[level][1])
            # keep the points whose patch stays inside the frame under the blur; evaluate on exactly those
            lv = prob.levels[level]
            m = 24 >> level
            keep = (xy[:, 0] > m) & (xy[:, 0] < lv.W - 1 - m) & (xy[:, 1] > m) & (xy[:, 1] < lv.H - 1 - m)
            lv.xy, lv.z = np.ascontiguousarray(xy[keep]), np.ascontiguousarray(z[keep])
    check_parity(O, gpu_eval(pkg, api, prob, 0), orc.evaluate(prob, 0))
    check_parity(O, gpu_eval(pkg, api, prob, 1), orc.evaluate(prob, 1))


def test_tma_staged_cost_pass_variant(pkg, api, O, orc, synth):
    """mbavo_debug_cost_tma — the cost pass with TMA-staged 8-bit keyframe tiles (cp.async.bulk.tensor.2d + mbarriers,
    csrc/track_cost_tma.cu; measured against the product pass in profiles/r2_tma_variant.md): the oracle's cost whatever the box,
    because samples outside a point's tile fall back to the global gather."""
    for name, level in (("tiny", 0), ("tiny", 1), ("C1", 0)):
        prob = synth.make_config(name)
        want = orc.evaluate(prob, level, with_hessian=False)[0]
        with pkg.Context(api.limits_for(prob)) as ctx:
            api.upload_problem(ctx, prob)
            a = (level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
            product = ctx.evaluate(*a, False)[0]
            fractions = []
            for box in ((16, 8), (32, 16), (64, 32)):
                for _ in range(2):  # the second call re-uses the mbarriers' phases from a fresh launch
                    c, ms, frac = ctx.cost_tma(*a, *box)
                assert abs(c - want) <= COST_TOL * want and abs(c - product) <= 1e-6 * product, (name, box, c, product)
                assert ms > 0
                fractions.append(frac)
            assert fractions[0] < fractions[-1] and fractions[-1] > 0.99, fractions  # small boxes really exercise the fallback
            with pytest.raises(pkg.MbavoError):
                ctx.cost_tma(*a, 24, 16)  # box width must be a multiple of 16


def test_keyframe_texels_and_direct_gather(pkg, api, O, orc, synth, monkeypatch):
    """mbavo_set_level packs the keyframe into patch texels (the 4 x 4 byte neighbourhood of every pixel) when byte differences
    reproduce every gradient value exactly (always for Gradient.h's central differences); the kernels then read bit-identical
    values through one 128-bit load per sample.  A gradient image the texels cannot hold makes the same kernels gather
    ref_I / ref_dIxy directly."""
    prob = synth.make_config("C1")
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        assert ctx.level_uses_texels(0) == 1
        a = gpu_eval(pkg, api, prob, ctx=ctx)
        a_cost = gpu_eval(pkg, api, prob, ctx=ctx, with_hessian=False)[0]
    monkeypatch.setenv("MBAVO_NO_TEXELS", "1")
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        assert ctx.level_uses_texels(0) == 0
        b = gpu_eval(pkg, api, prob, ctx=ctx)
        b_cost = gpu_eval(pkg, api, prob, ctx=ctx, with_hessian=False)[0]
    monkeypatch.delenv("MBAVO_NO_TEXELS")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert abs(a_cost - b_cost) <= 1e-7 * b_cost  # same tap values; only the FMA contraction of the blend may differ
    # a gradient image the texels cannot hold: scaled by 1/3
    prob.levels[0].ref_dIxy = (prob.levels[0].ref_dIxy / np.float32(3.0)).astype(np.float32)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        assert ctx.level_uses_texels(0) == 0
        check_parity(O, gpu_eval(pkg, api, prob, ctx=ctx), orc.evaluate(prob, 0))


@pytest.mark.parametrize("phases", [2, 4, 8, 32])
def test_exposure_phase_split(pkg, api, O, orc, synth, monkeypatch, phases):
    """MBAVO_PHASES: the lanes of a warp split into exposure phases x pixel slots (partial sums combined by warp shuffles)."""
    monkeypatch.setenv("MBAVO_PHASES", str(phases))
    for name in ("tiny", "C5cubic"):
        prob = synth.make_config(name)
        check_parity(O, gpu_eval(pkg, api, prob), orc.evaluate(prob, 0))
        c2 = gpu_eval(pkg, api, prob, with_hessian=False)[0]
        assert abs(c2 - orc.evaluate(prob, 0, with_hessian=False)[0]) <= COST_TOL * c2


def test_new_live_frame_keeps_the_keyframe(pkg, api, O, orc, synth):
    """mbavo_set_live_images: only the blurred frame changes between frames (tracker.cpp:112-116)."""
    prob = synth.make_config("tiny")
    other = synth.make_config("tiny")
    other.levels[0].cur_I = [np.ascontiguousarray(np.roll(c, 3, axis=1)) for c in prob.levels[0].cur_I]
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        first = gpu_eval(pkg, api, prob, ctx=ctx)
        ctx.set_live_images(0, other.levels[0].cur_I)
        check_parity(O, gpu_eval(pkg, api, other, ctx=ctx), orc.evaluate(other, 0))
        ctx.set_live_images(0, prob.levels[0].cur_I)
        again = gpu_eval(pkg, api, prob, ctx=ctx)
    assert again[0] == first[0] and np.array_equal(again[1], first[1])


@pytest.mark.parametrize("no_texels", [False, True])
def test_device_built_pyramid(pkg, api, O, orc, synth, monkeypatch, no_texels):
    """mbavo_set_keyframe_pyramid / set_live_pyramid / set_level_points: 2x2-box pyramid (ImagePyramid.h:59-99), central
    gradients (Gradient.h:17-75) and texels built on the GPU from the level-0 images give, on EVERY level, bit-identical
    results to the levels built on the CPU by the restatement of the same loops (synth.pyramid_down / image_gradient,
    pinned to the oracle's C restatement in tests/test_oracle.py)."""
    if no_texels:
        monkeypatch.setenv("MBAVO_NO_TEXELS", "1")
    prob = synth.make_problem("pyr", W=322, H=246, levels=4, P0=1500, N=8, n_knots=2, k=2, seed=5, margin=24)  # odd sizes on the way down
    with pkg.Context(api.limits_for(prob)) as a, pkg.Context(api.limits_for(prob)) as b:
        api.upload_problem(a, prob)
        api.upload_problem_pyramid(b, prob)
        for level in range(len(prob.levels)):
            assert b.level_uses_texels(level) == (0 if no_texels else 1)
            ra, rb = gpu_eval(pkg, api, prob, level, ctx=a), gpu_eval(pkg, api, prob, level, ctx=b)
            assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and np.array_equal(ra[3], rb[3])
            ca, cb = gpu_eval(pkg, api, prob, level, ctx=a, with_hessian=False), gpu_eval(pkg, api, prob, level, ctx=b, with_hessian=False)
            assert ca[0] == cb[0]
        check_parity(O, gpu_eval(pkg, api, prob, 3, ctx=b), orc.evaluate(prob, 3))
        # a new live frame: only its level 0 is uploaded
        other = [np.ascontiguousarray(np.roll(c, 2, axis=0)) for c in prob.levels[0].cur_I]
        b.set_live_pyramid(len(prob.levels), other)
        cur = other
        for level, lv in enumerate(prob.levels):
            if level > 0:
                cur = [synth.pyramid_down(c) for c in cur]
            lv.cur_I = [np.ascontiguousarray(c) for c in cur]
        check_parity(O, gpu_eval(pkg, api, prob, 2, ctx=b), orc.evaluate(prob, 2))


@pytest.mark.parametrize("async_upload", [False, True])
def test_set_frame_single_call_upload(pkg, api, O, orc, synth, async_upload):
    """mbavo_set_frame (keyframe + live frame + points of every level in one call, point copies on a second stream under the
    pyramid build, optionally without synchronisation) leaves exactly the state the three separate calls leave: bit-identical
    evaluations on every level; a later call with only a new live frame keeps keyframe and points; a sweep launched straight
    behind an asynchronous upload (its kernel waits, level by level, for the point copies) equals the sweep on the synchronously
    uploaded frame."""
    prob = synth.make_problem("frame", W=322, H=246, levels=4, P0=1500, N=8, n_knots=2, k=2, seed=6, margin=24)
    n = len(prob.levels)
    with pkg.Context(api.limits_for(prob)) as a, pkg.Context(api.limits_for(prob)) as b:
        api.upload_problem_pyramid(a, prob)
        b.set_frame_times(prob.cap, prob.exp)
        for _ in range(2):  # the second round re-uses every buffer
            b.set_frame(n, prob.levels[0].ref_I, prob.levels[0].cur_I, prob.levels, async_upload=async_upload)
            for level in range(n):
                ra, rb = gpu_eval(pkg, api, prob, level, ctx=a), gpu_eval(pkg, api, prob, level, ctx=b)
                assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and np.array_equal(ra[3], rb[3])
        other = [np.ascontiguousarray(np.roll(c, 2, axis=0)) for c in prob.levels[0].cur_I]
        a.set_live_pyramid(n, other)
        b.set_frame(n, None, other, prob.levels, async_upload=async_upload)
        for level in (0, n - 1):
            ra, rb = gpu_eval(pkg, api, prob, level, ctx=a), gpu_eval(pkg, api, prob, level, ctx=b)
            assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1])
        costs_a, kta, kRa = a.gn_sweep(n - 1, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        costs_b, ktb, kRb = b.gn_sweep(n - 1, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        assert np.array_equal(costs_a, costs_b) and np.array_equal(kta, ktb) and np.array_equal(kRa, kRb)
        # the sweep as the FIRST call behind the upload
        for _ in range(3):
            before = b.persistent_sweeps()
            b.set_frame(n, prob.levels[0].ref_I, other, prob.levels, async_upload=async_upload)
            costs_b, ktb, kRb = b.gn_sweep(n - 1, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
            assert b.persistent_sweeps() == before + 1
            assert np.array_equal(costs_a, costs_b) and np.array_equal(kta, ktb) and np.array_equal(kRa, kRb)
        # ... and an upload followed by something that is not a sweep
        b.set_frame(n, prob.levels[0].ref_I, other, prob.levels, async_upload=async_upload)
        ra, rb = gpu_eval(pkg, api, prob, 0, ctx=a), gpu_eval(pkg, api, prob, 0, ctx=b)
        assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1])


def test_synthetic_blurred_frame_bit_exact(pkg, api, orc, synth):
    """mbavo_synthesize_blurred (generate_synthetic_data.cpp:127-180) is byte work: bit-exact against the oracle's
    restatement on the same poses, and bit-exact against the reference-generated golden frame."""
    prob = synth.make_problem("blur", W=160, H=120, levels=1, P0=4, N=4, n_knots=3, k=2, seed=5, margin=10, motion_scale=3.0)
    I = prob.levels[0].ref_I
    ts = 0.1 + np.arange(24) * 1.7 / 23
    poses = np.array([np.concatenate(synth.spline_pose(2, prob.gt_knots_t, prob.gt_knots_R, 0.0, 1.0, t)) for t in ts])
    got = api.synthesize_blurred(I, 7.5, 80.0, 80.0, 80.0, 60.0, poses)
    want = orc.warp_mean(I, 7.5, 80.0, 80.0, 80.0, 60.0, poses)
    assert np.array_equal(got, want)
    assert (got == 0).any() and (got > 0).any()  # the strong motion leaves part of the frame without keyframe coverage
    # the frame the reference's own warp_image + synthesize_motion_blurred_img rendered (oracle/_ref), with the very poses its
    # spline functors gave for the exposure samples: byte for byte
    z = golden("blurred.npz")
    got = api.synthesize_blurred(np.ascontiguousarray(z["ref_I"]), float(z["D"]), float(z["fx"]), float(z["fy"]), float(z["cx"]),
                                 float(z["cy"]), np.ascontiguousarray(z["poses_tq"]))
    assert np.array_equal(got, z["out"])


def test_point_selection_bit_exact(pkg, api, O, orc, synth):
    """mbavo_select_points (FeatureDetectorSemiDense::detect, FeatureDetectorBase::gridSelection, the depth look-up of
    tmpProcessKeyframe) is byte / index work: the same points, in the same order, as the oracle — on the golden cases the
    reference's own detector produced, on VGA pyramids with the tracker's parameters, on a flat image (no points) — and the
    selected points are the level's points: an evaluation on them equals one on the same points uploaded from the host."""
    pattern = synth.make_config("C1").levels[0].pattern
    z = golden("point_selection.npz")
    for name in ("tex", "ramp"):
        I, depth = np.ascontiguousarray(z[name + "_I"]), np.ascontiguousarray(z[name + "_depth"])
        cnt = z[name + "_count"]
        offs = np.concatenate([[0], np.cumsum(cnt)])
        lim = api.Limits(max_num_keypoints=4096, max_num_virtual_poses_per_frame=8, max_patch_size=len(pattern))
        with pkg.Context(lim) as ctx:
            ctx.set_keyframe_pyramid(len(cnt), I)
            got_cnt = ctx.select_points(len(cnt), depth, 80.0, 80.0, 80.0, 60.0, pattern, 4, float(z[name + "_thr"]), int(z[name + "_cell"]),
                                        int(z[name + "_cell"]))
            assert got_cnt == list(cnt)
            for l in range(len(cnt)):
                xy, zz = ctx.get_points(l)
                assert np.array_equal(xy, z[name + "_xy"][offs[l]:offs[l + 1]]) and np.array_equal(zz, z[name + "_z"][offs[l]:offs[l + 1]])
        with pkg.Context(api.Limits(max_num_keypoints=50, max_num_virtual_poses_per_frame=8, max_patch_size=len(pattern))) as ctx:
            ctx.set_keyframe_pyramid(len(cnt), I)
            with pytest.raises(pkg.MbavoError):  # more points than max_num_keypoints
                ctx.select_points(len(cnt), depth, 80.0, 80.0, 80.0, 60.0, pattern, 4, float(z[name + "_thr"]), int(z[name + "_cell"]),
                                  int(z[name + "_cell"]))
            with pytest.raises(pkg.MbavoError):  # grid selection off is not the tracker's mode
                ctx.select_points(len(cnt), depth, 80.0, 80.0, 80.0, 60.0, pattern, 4, 25.0, -1, -1)
    rng = np.random.default_rng(11)
    prob = synth.make_config("C2")
    H, W = prob.levels[0].H, prob.levels[0].W
    depth = np.full((H, W), 7.5, np.float32)
    depth[rng.random((H, W)) < 0.05] = 0.0
    images = {"tex": prob.levels[0].ref_I, "ramp": synth.ramp_image(H, W), "noise": rng.integers(0, 256, (H, W)).astype(np.uint8),
              "flat": np.full((H, W), 9, np.uint8)}
    lim = api.limits_for(prob)
    with pkg.Context(lim) as ctx:
        for name, I in images.items():
            for thr, ch, cw in ((25.0, 30, 30), (2.0, 7, 5), (0.5, 17, 23)):
                ctx.set_keyframe_pyramid(4, I)
                want = O.select_points(I, 4, thr, ch, cw, depth)
                cnt = ctx.select_points(4, depth, 320.0, 320.0, 320.0, 240.0, pattern, 16, thr, ch, cw)
                assert cnt == [len(zz) for _, zz in want], (name, thr)
                for l in range(4):
                    xy, zz = ctx.get_points(l)
                    assert np.array_equal(xy, want[l][0]) and np.array_equal(zz, want[l][1]), (name, thr, l)
                if name == "flat":
                    assert cnt == [0, 0, 0, 0]
        # the selected points drive the tracker: same result as the same points set from the host
        ctx.set_frame_times(prob.cap, prob.exp)
        ctx.set_keyframe_pyramid(4, prob.levels[0].ref_I)
        ctx.set_live_pyramid(4, prob.levels[0].cur_I)
        cnt = ctx.select_points(4, depth, prob.levels[0].fx, prob.levels[0].fy, prob.levels[0].cx, prob.levels[0].cy, pattern,
                                prob.levels[0].N, 2.0, 7, 5)
        want = O.select_points(prob.levels[0].ref_I, 4, 2.0, 7, 5, depth)
        with pkg.Context(lim) as other:
            api.upload_problem_pyramid(other, prob)
            for l in (0, 2, 3):
                lv = prob.levels[l]
                lv.xy, lv.z = np.ascontiguousarray(want[l][0]), np.ascontiguousarray(want[l][1])
                other.set_level_points(l, lv)
                a = ctx.evaluate(l, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
                b = other.evaluate(l, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
                assert cnt[l] > 500 and a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_keyframe_statistics(pkg, api, O, synth):
    """mbavo_keyframe_stats (isKeyframe, tracker.cpp:205-248): mean flow and mean blur-kernel length of the host-map points."""
    prob = synth.make_config("C1")
    lv = prob.levels[0]
    cap, exp = float(prob.cap[0]), float(prob.exp[0])
    poses = np.array([np.concatenate(synth.spline_pose(prob.k, prob.gt_knots_t, prob.gt_knots_R, prob.t0, prob.dt, t))
                      for t in (cap, cap - 0.5 * exp, cap + 0.5 * exp)])
    want = O.keyframe_stats(lv.xy, lv.z, lv.fx, lv.fy, lv.cx, lv.cy, poses)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        got = ctx.keyframe_stats(0, poses)
    assert want[0] > 0.1 and want[1] > 0.1
    assert abs(got[0] - want[0]) <= 1e-6 * want[0] and abs(got[1] - want[1]) <= 1e-6 * want[1]  # float sqrt of fp64 sums


def test_multiple_frames(pkg, api, O, orc, synth):
    """n_frames > 1 with frames in different segments (the merge of overlapping frames, test_merge…:1060-1181)."""
    prob = synth.make_problem("frames", W=160, H=120, levels=1, P0=300, N=8, n_knots=3, k=2, seed=8, margin=14, F=2)
    prob.dt = 1.0
    prob.cap, prob.exp = np.array([0.5, 1.45]), np.array([0.8, 0.8])
    check_parity(O, gpu_eval(pkg, api, prob), orc.evaluate(prob, 0))


def test_outlier_flags_and_device_side_detection(pkg, api, O, orc, synth):
    prob = synth.make_config("tiny")
    lv = prob.levels[0]
    flags = np.zeros(lv.P, dtype=np.uint8)
    flags[5::9] = 1
    nbad = int(flags.sum())
    check_parity(O, gpu_eval(pkg, api, prob, flags=flags, num_bad=nbad), orc.evaluate(prob, 0, flags=flags, num_bad=nbad))
    # detectOutliersAndUploadToGpu on the device == the restatement applied to the same patch costs
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        kt = prob.knots_t + 0.05  # a bad pose so that there is a spread of patch costs
        c, _, _ = ctx.evaluate(0, prob.k, prob.t0, prob.dt, kt, prob.knots_R, prob.huber_a, False)
        pc = ctx.patch_costs(0, prob.F, lv.P)
        for k_sigma in (3.0, 1.0):
            want_flags = np.zeros(lv.P, dtype=np.uint8)
            want_n = O.detect_outliers(pc, want_flags, k_sigma)
            ctx.set_outliers(0, None)
            got_n = ctx.detect_outliers(0, k_sigma)
            assert got_n == want_n and want_n > 0
            # the flags took effect: the next evaluation equals the oracle's with the same flags
            got = ctx.evaluate(0, prob.k, prob.t0, prob.dt, kt, prob.knots_R, prob.huber_a, True)
            want = orc.evaluate(prob, 0, kt, prob.knots_R, flags=want_flags, num_bad=want_n)
            assert abs(got[0] - want[0]) <= COST_TOL * want[0]
            assert max_rel(got[1], want[1]) <= 1e-4


@pytest.mark.parametrize("name", ["tiny", "C2"])
def test_lm_loop_matches_oracle(pkg, api, O, orc, synth, name):
    """optimizePyramidLevel (tracker.cpp:590-637) coarse to fine: same accept / reject sequence, same final knots."""
    prob = synth.make_config(name)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        kt, kR, summ = pkg.optimize_trajectory(ctx, prob)
    kt_o, kR_o, traces = O.optimize_trajectory(orc, prob)
    for s, tr in zip(summ, traces):
        # the very first step of every level is the parity quantity
        assert rel(s["first_step"], tr.first_step) <= DELTA_TOL
        same = s["decisions"] == "".join(tr.decisions)
        if same:
            assert abs(s["final_cost"] - tr.costs[-1]) <= 1e-4 * tr.costs[-1]
            assert s["num_bad_keypoints"] == tr.num_bad
        else:  # a decision may differ only at a near-tie; then the final cost decides
            assert abs(s["final_cost"] - tr.costs[-1]) <= 1e-3 * tr.costs[-1]
    assert np.abs(kt - kt_o).max() <= 1e-4 and np.abs(kR - kR_o).max() <= 1e-5
    if name == "C2":  # converged towards the ground-truth trajectory (plane scene)
        assert np.abs(kt - prob.gt_knots_t).max() < np.abs(prob.knots_t - prob.gt_knots_t).max() * 5 + 0.02


def test_deterministic(pkg, api, synth):
    """Fixed-order reductions: bit-identical results run to run and context to context."""
    prob = synth.make_config("C1")
    a = gpu_eval(pkg, api, prob)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        runs = [gpu_eval(pkg, api, prob, ctx=ctx) for _ in range(3)]
    for r in runs:
        assert r[0] == a[0] and np.array_equal(r[1], a[1]) and np.array_equal(r[2], a[2]) and np.array_equal(r[3], a[3])
    b = gpu_eval(pkg, api, prob)
    assert b[0] == a[0] and np.array_equal(b[1], a[1]) and np.array_equal(b[2], a[2])


def test_error_codes(pkg, api, synth):
    prob = synth.make_config("tiny")
    with pkg.Context(api.limits_for(prob)) as ctx:
        with pytest.raises(pkg.MbavoError, match="-5"):  # MBAVO_ENOTREADY: level not set
            ctx.evaluate(0, 2, 0.0, 1.0, prob.knots_t, prob.knots_R, 10.0)
        api.upload_problem(ctx, prob)
        with pytest.raises(pkg.MbavoError, match="-3"):  # MBAVO_ERANGE: exposure window leaves the spline
            ctx.evaluate(0, 2, 0.0, 0.5, prob.knots_t, prob.knots_R, 10.0)
        with pytest.raises(pkg.MbavoError, match="-3"):
            ctx.evaluate(0, 2, 2.0, 1.0, prob.knots_t, prob.knots_R, 10.0)
        with pytest.raises(pkg.MbavoError, match="-1"):  # MBAVO_EINVAL: spline order
            ctx.evaluate(0, 3, 0.0, 1.0, prob.knots_t, prob.knots_R, 10.0)
        big = synth.make_problem("big", W=64, H=48, levels=1, P0=prob.levels[0].P + 1, N=4, n_knots=2, seed=1, margin=10)
        with pytest.raises(pkg.MbavoError, match="-4"):  # MBAVO_ECAPACITY
            ctx.set_level(0, big.levels[0])
        # the context is still usable after errors
        assert ctx.evaluate(0, 2, prob.t0, prob.dt, prob.knots_t, prob.knots_R, 10.0)[0] > 0


def test_full_size_properties(pkg, api, O, synth):
    """BASELINE C3 at full size (1280x720, 80k points, 32 samples, 3 knots) through size-independent properties:
    cost-only pass == cost of the Hessian pass; H symmetric positive semi-definite; linearity: the packed vectors of two
    point shards (global normaliser) sum to the unsharded result; g.delta < 0 (descent direction)."""
    import torch

    prob = synth.make_config("C3", levels=1)
    lv = prob.levels[0]
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        c, H, g = ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
        c2, _, _ = ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
        pc = ctx.patch_costs(0, 1, lv.P)
    assert abs(c - c2) <= 1e-6 * c                # the two passes blend the same tap values in different FMA order
    assert abs(pc.sum() - c) <= 1e-9 * c          # checksum of checksums: patch costs add up to the total
    assert np.array_equal(H, H.T) and np.linalg.eigvalsh(H).min() >= -1e-9 * np.abs(H).max()
    d = first_step(O, H, g)
    assert g @ d < 0
    total = None
    half = (lv.P // 2 // 32) * 32 + 32  # a multiple of the warp batch: the fp32 partial sums of both shards are the unsharded ones
    nres = lv.P * prob.F * lv.S
    for sl in (slice(0, half), slice(half, lv.P)):
        with pkg.Context(api.limits_for(prob)) as ctx:
            ctx.set_frame_times(prob.cap, prob.exp)
            ctx.set_level(0, lv, sl)
            buf = torch.zeros(ctx.packed_len(8), dtype=torch.float64, device="cuda")
            kmin, nk = ctx.evaluate_async(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True, nres,
                                          buf.data_ptr())
            torch.cuda.synchronize()
            v = buf[: ctx.packed_len(nk)].cpu().numpy()
            total = v if total is None else total + v
            unpack = ctx.unpack
            cs, Hs, gs = unpack(total, kmin, nk, prob.n_knots)
    assert abs(cs - c) <= 1e-12 * c and max_rel(Hs, H) <= 1e-12 and max_rel(gs, g) <= 1e-12


def test_sharded_evaluator_single_rank(pkg, api, synth):
    """mbavo_b200.parallel.ShardedEvaluator with world_size 1 on a torch stream == the blocking C-ABI call."""
    import torch

    from mbavo_b200.parallel import ShardedEvaluator

    prob = synth.make_config("tiny")
    want = gpu_eval(pkg, api, prob)
    torch.cuda.set_device(0)
    with pkg.Context(api.limits_for(prob)) as ctx:
        ev = ShardedEvaluator(ctx, prob, 0, 1)
        c, H, g = ev.evaluate(0, prob.knots_t, prob.knots_R, True)
        c2, _, _ = ev.evaluate(0, prob.knots_t, prob.knots_R, False)
    assert c == want[0] and np.array_equal(H, want[1]) and np.array_equal(g, want[2])
    assert abs(c2 - c) <= 1e-6 * c  # the cost-only pass and the Hessian pass read the same tap values from different texels


@pytest.mark.parametrize("world", [2, 3])
def test_fused_shard_allreduce(pkg, api, O, orc, synth, world):
    """mbavo_shard_*: `world` ranks (contexts on this GPU, one thread each, mailboxes connected by pointer) evaluate their
    point shards; the kernels exchange the packed vectors through the mailboxes and every rank returns the global result.
    Checked against the unsharded context and the oracle; then the collective outlier detection and a collective LM loop."""
    from mbavo_b200.parallel import shard_bounds

    prob = synth.make_config("tiny")
    lv = prob.levels[0]
    ctxs = [pkg.Context(api.limits_for(prob)) for _ in range(world)]
    try:
        ptrs = []
        for r, ctx in enumerate(ctxs):
            ctx.set_frame_times(prob.cap, prob.exp)
            for l, lvl in enumerate(prob.levels):
                lo, hi = shard_bounds(lvl.P, r, world)
                ctx.set_level(l, lvl, slice(lo, hi))
            ptrs.append(ctx.shard_export()[1])
        for r, ctx in enumerate(ctxs):
            ctx.shard_connect(world, r, mailbox_ptrs=ptrs)
            for l, lvl in enumerate(prob.levels):
                ctx.shard_set_global_points(l, lvl.P)
        args = (0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
        res = _run_ranks([lambda c=c: c.evaluate(*args, True) for c in ctxs])
        res_c = _run_ranks([lambda c=c: c.evaluate(*args, False) for c in ctxs])
        want = orc.evaluate(prob, 0)
        for (c, H, g), (c2, _, _) in zip(res, res_c):
            assert c == res[0][0] and np.array_equal(H, res[0][1]) and np.array_equal(g, res[0][2])  # identical on all ranks
            assert abs(c - want[0]) <= COST_TOL * want[0] and abs(c2 - want[0]) <= COST_TOL * want[0]
            assert max_rel(H, want[1]) <= 1e-5 and max_rel(g, want[2]) <= 1e-5
            assert rel(first_step(O, H, g), first_step(O, want[1], want[2])) <= DELTA_TOL
        # collective outlier statistics == the restatement applied to ALL patch costs
        kt = prob.knots_t + 0.05
        _run_ranks([lambda c=c: c.evaluate(0, prob.k, prob.t0, prob.dt, kt, prob.knots_R, prob.huber_a, False) for c in ctxs])
        pcs = [ctx.patch_costs(0, prob.F, shard_bounds(lv.P, r, world)[1] - shard_bounds(lv.P, r, world)[0])[0]
               for r, ctx in enumerate(ctxs)]
        want_flags = np.zeros(lv.P, dtype=np.uint8)
        want_n = O.detect_outliers(np.concatenate(pcs)[None, :], want_flags, 3.0)
        got_n = _run_ranks([lambda c=c: c.detect_outliers(0, 3.0) for c in ctxs])
        assert want_n > 0 and all(n == want_n for n in got_n)
        got = _run_ranks([lambda c=c: c.evaluate(0, prob.k, prob.t0, prob.dt, kt, prob.knots_R, prob.huber_a, True) for c in ctxs])
        want2 = orc.evaluate(prob, 0, kt, prob.knots_R, flags=want_flags, num_bad=want_n)
        assert abs(got[0][0] - want2[0]) <= COST_TOL * want2[0] and max_rel(got[0][1], want2[1]) <= 1e-4
        # collective keyframe statistics: means over the points of all ranks
        cap, exp = float(prob.cap[0]), float(prob.exp[0])
        poses = np.array([np.concatenate(synth.spline_pose(prob.k, prob.gt_knots_t, prob.gt_knots_R, prob.t0, prob.dt, t))
                          for t in (cap, cap - 0.5 * exp, cap + 0.5 * exp)])
        kf = _run_ranks([lambda c=c: c.keyframe_stats(0, poses) for c in ctxs])
        kf_want = O.keyframe_stats(lv.xy, lv.z, lv.fx, lv.fy, lv.cx, lv.cy, poses)
        assert all(k == kf[0] for k in kf) and abs(kf[0][0] - kf_want[0]) <= 1e-6 * kf_want[0] and abs(kf[0][1] - kf_want[1]) <= 1e-6 * kf_want[1]
        # a whole collective LM loop: every rank commits the same knots, equal to the single-context run
        for c in ctxs:
            c.set_outliers(0, None)
        lm = _run_ranks([lambda c=c: pkg.optimize_trajectory(c, prob) for c in ctxs])
        with pkg.Context(api.limits_for(prob)) as solo:
            api.upload_problem(solo, prob)
            kt1, kR1, s1 = pkg.optimize_trajectory(solo, prob)
        for kt_r, kR_r, s_r in lm:
            assert np.array_equal(kt_r, lm[0][0]) and np.array_equal(kR_r, lm[0][1])
            # (the sharded sums add the same terms in another order; the LM iterations amplify that rounding)
            assert np.abs(kt_r - kt1).max() <= 1e-4 and np.abs(kR_r - kR1).max() <= 1e-5
            assert [x["decisions"] for x in s_r] == [x["decisions"] for x in s1]
    finally:
        for c in ctxs:
            c.close()


def test_shard_peer_timeout_is_an_error(pkg, api, synth):
    """A rank whose peer never calls gets MBAVO_ENCCL after the in-kernel time-out instead of hanging the GPU."""
    prob = synth.make_config("tiny")
    a, b = pkg.Context(api.limits_for(prob)), pkg.Context(api.limits_for(prob))
    try:
        ptrs = []
        for ctx in (a, b):
            api.upload_problem(ctx, prob)
            ptrs.append(ctx.shard_export()[1])
        a.shard_connect(2, 0, mailbox_ptrs=ptrs)
        a.shard_set_global_points(0, 2 * prob.levels[0].P)
        with pytest.raises(pkg.MbavoError, match="-6"):
            a.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
        a.shard_disconnect()
        assert a.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)[0] > 0
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("name,chain", [("tiny", False), ("C2", True), ("C2", False), ("C5cubic", True)])
def test_device_resident_sweep_matches_host_path(pkg, api, synth, monkeypatch, name, chain):
    """mbavo_gn_sweep in its three forms: ONE persistent launch (sweep_kernel: every pass of every level inside one resident
    grid, the candidate's sample records computed by the pass's last block), the per-pass device-resident form
    (MBAVO_NO_PERSISTENT=1: solve, candidate and commit in the kernels' last blocks, one host wait) and the
    evaluation-by-evaluation form (MBAVO_NO_DEVICE_SWEEP=1)."""
    prob = synth.make_config(name)
    top = len(prob.levels) - 1

    def sweep():
        with pkg.Context(api.limits_for(prob)) as ctx:
            api.upload_problem(ctx, prob)
            l0 = ctx.kernel_launches()
            out = [ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=chain)
                   for _ in range(3)]
            for o in out[1:]:
                assert np.array_equal(out[0][0], o[0]) and np.array_equal(out[0][1], o[1])  # deterministic
            return out[0] + (ctx.device_sweeps(), ctx.persistent_sweeps(), ctx.kernel_launches() - l0, ctx.lib.mbavo_last_error().decode())

    pers = sweep()
    monkeypatch.setenv("MBAVO_NO_PERSISTENT", "1")
    dev = sweep()
    monkeypatch.setenv("MBAVO_NO_DEVICE_SWEEP", "1")
    host = sweep()
    for got in (pers, dev):
        assert np.abs(got[0] - host[0]).max() <= 1e-9 * np.abs(host[0]).max(), (got[0], host[0])
        assert np.abs(got[1] - host[1]).max() <= 1e-9 and np.abs(got[2] - host[2]).max() <= 1e-9
    if chain:
        assert np.abs(dev[1] - prob.knots_t).max() > 0  # a candidate was committed
    if name == "C5cubic":  # a 7-knot window has no persistent instantiation: that sweep runs pass by pass
        assert pers[3] == 3 and pers[4] == 0, pers[3:]
    else:
        assert pers[3] == 3 and pers[4] == 3 and pers[5] == 6, pers[3:]  # two launches per sweep: pose kernel + sweep kernel
    assert dev[3] == 3 and dev[4] == 0 and host[3] == 0, (dev[3:], host[3:])  # the device-resident paths really ran


@pytest.mark.parametrize("seed,mixed_n", [(2, False), (3, False), (3, True)])
def test_device_sweep_rejected_levels_and_record_reuse(pkg, api, synth, monkeypatch, seed, mixed_n):
    """A chained sweep in which some level's candidate is REJECTED (the sweep keeps standing on its knots) and others are
    committed: the levels after the first find the sample records of their knots in one of the two record buffers instead of
    running a pose kernel — whichever buffer that is after the commits so far.  With a different number of exposure samples on
    one level the records cannot be shared: every level computes its own, one launch per pass.  The persistent launch, the
    per-pass form and the evaluation-by-evaluation form against each other."""
    prob = synth.make_problem("rej", W=160, H=120, levels=3, P0=600, N=8, n_knots=2, k=2, seed=100 + seed, margin=16)
    kt = prob.knots_t + np.random.default_rng(seed).normal(size=prob.knots_t.shape) * 2e-2
    if mixed_n:
        prob.levels[1].N = 5

    def sweep():
        with pkg.Context(api.limits_for(prob)) as ctx:
            api.upload_problem(ctx, prob)
            l0 = ctx.kernel_launches()
            out = ctx.gn_sweep(2, 0, prob.k, prob.t0, prob.dt, kt, prob.knots_R, prob.huber_a, 1e4, chain=True)
            return out + (ctx.device_sweeps(), ctx.kernel_launches() - l0, ctx.persistent_sweeps())

    pers = sweep()
    monkeypatch.setenv("MBAVO_NO_PERSISTENT", "1")
    dev = sweep()
    monkeypatch.setenv("MBAVO_NO_DEVICE_SWEEP", "1")
    host = sweep()
    assert dev[3] == 1 and host[3] == 0 and pers[3] == 1 and pers[5] == (0 if mixed_n else 1)
    assert dev[4] == (12 if mixed_n else 10) and host[4] == 12  # 6 tracking kernels + 6 pose kernels, or 4 with shared records
    assert pers[4] == (12 if mixed_n else 2)
    for got in (pers, dev):
        assert np.abs(got[0] - host[0]).max() <= 1e-9 * np.abs(host[0]).max(), (got[0], host[0])
        assert np.abs(got[1] - host[1]).max() <= 1e-9 and np.abs(got[2] - host[2]).max() <= 1e-9
        decisions = "".join("A" if cand < cost else "R" for cost, cand in got[0])
        if not mixed_n:
            assert decisions == {2: "AAR", 3: "RAA"}[seed], decisions  # (predicted with the oracle on the CPU)


def test_device_sweep_with_unobserved_knots(pkg, api, O, orc, synth, monkeypatch):
    """A spline with a control knot that no exposure sample touches: the normal equations are singular in that knot.  The
    host path solves them by the pseudo-inverse (zero step on the unobserved knot); the device-resident sweep solves the
    window system only, which is the same step."""
    prob = synth.make_problem("unobs", W=192, H=144, levels=2, P0=900, N=8, n_knots=3, k=2, seed=21, margin=16)
    prob.dt = float(prob.exp[0]) * 1.0001  # the whole exposure lies in segment 0: knot 2 is unobserved
    prob.cap[:] = prob.t0 + 0.5 * prob.exp[0]

    def sweep():
        with pkg.Context(api.limits_for(prob)) as ctx:
            api.upload_problem(ctx, prob)
            out = ctx.gn_sweep(1, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
            return out + (ctx.device_sweeps(),)

    dev = sweep()
    monkeypatch.setenv("MBAVO_NO_DEVICE_SWEEP", "1")
    host = sweep()
    assert dev[3] == 1 and host[3] == 0
    assert np.abs(dev[0] - host[0]).max() <= 1e-8 * np.abs(host[0]).max()
    assert np.abs(dev[1] - host[1]).max() <= 1e-8 and np.abs(dev[2] - host[2]).max() <= 1e-8
    assert np.array_equal(dev[1][2], prob.knots_t[2]) and np.array_equal(dev[2][2], prob.knots_R[2])  # untouched
    want = orc.evaluate(prob, 1)
    assert abs(dev[0][0, 0] - want[0]) <= COST_TOL * want[0]
