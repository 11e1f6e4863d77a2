"""Generates tests/golden/*.npz from oracle/_ref — the REFERENCE's own header arithmetic
(compute_pixel_intensity.h, SplineFunctor.h, Quaternion.h, SmallBlas.h compiled from /root/reference, see
oracle/ref_harness.cpp) — so that machines without /root/reference can still pin the oracle and the CUDA path to
reference outputs.  The reference ships no golden vectors of its own.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
Inputs are stored next to the outputs; images are stored as uint8, gradients are recomputed by the consumer.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
from conftest import reference_test_spline  # noqa: E402
from oracle import oracle as O  # noqa: E402

pkg = ge.load_package()
synth = pkg.synth


def problem_arrays(prob):
    lv = prob.levels[0]
    return dict(ref_I=lv.ref_I, cur_I=np.stack(lv.cur_I), xy=lv.xy, z=lv.z, pattern=lv.pattern, N=lv.N, H=lv.H, W=lv.W,
                fx=lv.fx, fy=lv.fy, cx=lv.cx, cy=lv.cy, cap=prob.cap, exp=prob.exp, k=prob.k, t0=prob.t0, dt=prob.dt,
                knots_t=prob.knots_t, knots_R=prob.knots_R, huber_a=prob.huber_a, seg_start=prob.seg_start)


def golden_blurred(ref):
    """4. synthetic blurred frame (generate_synthetic_data.cpp:152-180), stored with the very poses the reference functors gave
    for its exposure samples (sample time as :161: capture - exposure / 2 + i * exposure / (n - 1))."""
    prob = synth.make_problem("golden_blur", W=96, H=64, levels=1, P0=4, N=4, n_knots=2, seed=5, margin=10)
    cap, exp, n = 0.45, 0.9, 16
    out = ref.synthesize_blurred(prob.levels[0].ref_I, 7.5, 48.0, 48.0, 48.0, 32.0, 2, 0.0, 1.0, prob.gt_knots_t,
                                 prob.gt_knots_R, cap, exp, n)
    poses = np.stack([ref.spline_pose(2, 0.0, 1.0, prob.gt_knots_t, prob.gt_knots_R, cap - exp * 0.5 + i * exp / (n - 1)) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, "blurred.npz"), ref_I=prob.levels[0].ref_I, knots_t=prob.gt_knots_t,
                        knots_R=prob.gt_knots_R, D=7.5, fx=48.0, fy=48.0, cx=48.0, cy=32.0, cap=cap, exp=exp, n=n, out=out, poses_tq=poses)


def golden_shapes():
    """7. the reference's synthetic scene (generate_synthetic_data.cpp:11-125: synthesize_img_with_rand_shapes) drawn by OpenCV's own
    cv::fillPoly through cv2 — same vertices, colours and LINE_8 as the reference passes."""
    import cv2

    im = np.zeros((480, 640), np.uint8)
    for (x, y), (w, h) in synth.SHAPES_RECTS:
        cv2.fillPoly(im, [np.array([[x, y], [x + w, y], [x + w, y + h], [x, y + h]], np.int32)], 255, cv2.LINE_8)
    for tri in synth.SHAPES_TRIANGLES:
        cv2.fillPoly(im, [np.array(tri, np.int32)], 255, cv2.LINE_8)
    np.savez_compressed(os.path.join(HERE, "shapes.npz"), image=im, opencv_version=cv2.__version__)


def main():
    if "--only-shapes" in sys.argv:
        golden_shapes()
        return
    ref = O.RefLib()
    if "--only-blurred" in sys.argv:
        golden_blurred(ref)
        return
    # 1. virtual poses + Jacobians on the reference test's 7-knot spline (test_compute_virtual_camera_poses, :183-342)
    kt, kR = reference_test_spline()
    cap = 0.25 + 0.5 * np.arange(4)
    exp = np.full(4, 0.1)
    for k in (2, 4):
        poses, seg, Jt, JR = ref.virtual_poses(32, cap, exp, k, 0.0, 0.5, kt, kR)
        np.savez_compressed(os.path.join(HERE, f"virtual_poses_k{k}.npz"), knots_t=kt, knots_R=kR, cap=cap, exp=exp, N=32,
                            t0=0.0, dt=0.5, poses=poses, seg=seg, Jt=Jt, JR=JR)

    # 2. compute_pixel_intensity on the reference ramp image (test_compute_pixel_intensity, :83-181)
    rng = np.random.default_rng(99)
    I = synth.ramp_image(120, 160)
    g = synth.image_gradient(I)
    rows = []
    for _ in range(64):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        q = np.array([0.05 * q[0], 0.05 * q[1], 0.05 * q[2], 1.0])
        q /= np.linalg.norm(q)
        pose = np.concatenate([rng.uniform(-0.3, 0.3, 3), q])
        D = rng.uniform(5, 10)
        X, Y = float(rng.integers(10, 150)), float(rng.integers(10, 110))
        ok, inten, J = ref.pixel_intensity(I, g, pose, D, 80.0, 80.0, 80.0, 60.0, X, Y)
        rows.append(np.concatenate([pose, [D, X, Y, float(ok), inten], J]))
    np.savez_compressed(os.path.join(HERE, "pixel_intensity.npz"), H=120, W=160, fx=80.0, fy=80.0, cx=80.0, cy=60.0,
                        rows=np.array(rows))

    # 3. whole evaluations (single segment, the reference's design point): k=2/n=2, k=4/n=4, and with outlier flags
    for tag, kw in (("k2", dict(k=2, n_knots=2)), ("k4", dict(k=4, n_knots=4))):
        prob = synth.make_problem(f"golden_{tag}", W=128, H=96, levels=1, P0=60, N=8, seed=4242, margin=14, **kw)
        arrs = problem_arrays(prob)
        c, Hm, gv, pc = ref.evaluate(prob, 0)
        c2, _, _, _ = ref.evaluate(prob, 0, with_hessian=False)
        flags = np.zeros(prob.levels[0].P, dtype=np.uint8)
        flags[::7] = 1
        cf, Hf, gf, pcf = ref.evaluate(prob, 0, flags=flags, num_bad=int(flags.sum()))
        np.savez_compressed(os.path.join(HERE, f"evaluate_{tag}.npz"), cost=c, Hessian=Hm, gradient=gv, patch_costs=pc, cost_only=c2,
                            flags=flags, cost_flagged=cf, H_flagged=Hf, g_flagged=gf, patch_costs_flagged=pcf, **arrs)

    golden_blurred(ref)  # 4.
    # 5. semi-dense point selection by the reference's own detector sources (oracle/_ref/libmbavo_refselect.so):
    #    a textured image and the reference test's ramp image (ties in every cell), a depth map with holes
    sel = O.RefSelect()
    rng = np.random.default_rng(7)
    cases = {}
    for name, img in (("tex", synth.make_problem("golden_sel", W=160, H=120, levels=1, P0=4, N=4, n_knots=2, seed=9, margin=10).levels[0].ref_I),
                      ("ramp", synth.ramp_image(120, 160))):
        depth = rng.uniform(0.5, 10.0, img.shape).astype(np.float32)
        depth[rng.random(img.shape) < 0.15] = 0.0
        thr, cell = (4.0, 12) if name == "tex" else (0.25, 9)
        res = sel.select_points(img, 3, thr, cell, cell, depth)
        cases[name + "_I"] = img
        cases[name + "_depth"] = depth
        cases[name + "_thr"] = thr
        cases[name + "_cell"] = cell
        cases[name + "_count"] = np.array([len(z) for _, z in res])
        cases[name + "_xy"] = np.concatenate([xy for xy, _ in res])
        cases[name + "_z"] = np.concatenate([z for _, z in res])
    np.savez_compressed(os.path.join(HERE, "point_selection.npz"), **cases)

    # 6. the reference's own pyramid and gradient loops (ImagePyramid.h:59-99, Gradient.h:17-75) on an odd-sized image
    img = np.random.default_rng(12).integers(0, 256, (47, 61)).astype(np.uint8)
    pyr = sel.pyramid(img, 3)
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **{f"I{l}": im for l, (im, _) in enumerate(pyr)},
                        **{f"g{l}": g for l, (_, g) in enumerate(pyr)})
    golden_shapes()  # 7.
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
