"""The REFERENCE's own CUDA kernels as a second oracle (run with -m gpu): src/ba_tracker/compute_virtual_camera_poses.cu,
compute_local_patches_xy.cu and compute_hessian_gradients_cost.cu, compiled unmodified for sm_100a into
oracle/_ref/libmbavo_refcuda.so (oracle/Makefile, built where /root/reference exists; the binary travels to the GPU box)
behind a restated orchestration (oracle/ref_cuda_harness.cu).  Their results on this B200 are held against

  * the committed golden vectors (tests/golden/evaluate_k{2,4}.npz: outputs of the reference's header arithmetic on the CPU),
  * the C restatement (oracle/libmbavo_oracle.so) — which therefore is pinned to the reference's GPU path too, and
  * the product library on the same inputs,

for linear (k = 2) and cubic (k = 4) splines, two frames, and outlier flags.  The reference kernels need power-of-two sample counts
and patch sizes (reduction.h:13-55) and attribute every sample's Jacobian to the frame's capture-time segment, so the cases
keep each exposure inside one segment, where that is what the oracle computes too.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import first_step, golden, max_rel, problem_from_golden, rel

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libmbavo_refcuda.so")
# reference kernels on the GPU vs the reference arithmetic / the oracle on the CPU: fp64 everywhere except the fp32 bilinear
# blend and the sqrtf of the Huber weight, whose FMA contraction differs between nvcc and the host compiler (measured <= 2e-7 with outlier flags, <= 5e-9 without)
REF_TOL = 1e-6
REF_TOL_G = 2e-6  # the gradient is a sum of signed terms that nearly cancel at these inputs (measured 2e-7 of its largest entry)


class RefCuda:
    def __init__(self, prob):
        self.lib = C.CDLL(LIB)
        self.lib.mbavo_refcuda_create.restype = C.c_void_p
        self.prob = prob
        maxP = max(lv.P for lv in prob.levels)
        h = self.lib.mbavo_refcuda_create(prob.F, 64, maxP, 8, 16, prob.k)
        assert h, "mbavo_refcuda_create failed"
        self.h = C.c_void_p(h)

    def set_level(self, level):
        p, lv = self.prob, self.prob.levels[level]
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        cur = (C.c_void_p * p.F)(*[c.ctypes.data for c in lv.cur_I])
        rc = self.lib.mbavo_refcuda_set_level(self.h, lv.H, lv.W, C.c_double(lv.fx), C.c_double(lv.fy), C.c_double(lv.cx), C.c_double(lv.cy),
                                              C.c_void_p(lv.ref_I.ctypes.data), C.c_void_p(lv.ref_dIxy.ctypes.data), cur, p.F,
                                              dp(np.ascontiguousarray(p.cap)), dp(np.ascontiguousarray(p.exp)), dp(lv.xy), dp(lv.z), lv.P,
                                              C.c_void_p(lv.pattern.ctypes.data), lv.S, lv.N)
        assert rc == 0, rc

    def evaluate(self, with_h=True, flags=None, num_bad=0):
        p = self.prob
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        n = p.n_knots
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        assert self.lib.mbavo_refcuda_set_flags(self.h, None if fl is None else fl.ctypes.data_as(C.POINTER(C.c_ubyte))) == 0
        H, g, cost = np.zeros((6 * n, 6 * n)), np.zeros(6 * n), C.c_double(0)
        kt, kR = np.ascontiguousarray(p.knots_t), np.ascontiguousarray(p.knots_R)
        seg = np.ascontiguousarray(p.seg_start, dtype=np.int32)
        rc = self.lib.mbavo_refcuda_evaluate(self.h, C.c_double(p.t0), C.c_double(p.dt), dp(kt), dp(kR), n, seg.ctypes.data_as(C.POINTER(C.c_int)),
                                             C.c_double(p.huber_a), C.c_int(num_bad), C.byref(cost), dp(H) if with_h else None,
                                             dp(g) if with_h else None)
        assert rc == 0, rc
        return cost.value, (H if with_h else None), (g if with_h else None)

    def close(self):
        self.lib.mbavo_refcuda_destroy(self.h)


@pytest.fixture(scope="module")
def api(pkg):
    from mbavo_b200 import api as a

    return a


def need_lib():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libmbavo_refcuda.so not built (no /root/reference where this tree was built)")


@pytest.mark.parametrize("tag", ["k2", "k4"])
def test_reference_cuda_kernels_reproduce_the_golden_vectors(pkg, api, O, orc, synth, tag):
    """Golden inputs -> the reference's CUDA kernels on this GPU: cost / H / g equal the committed outputs of the reference's CPU
    arithmetic to REF_TOL (summation order, FMA contraction of the fp32 blend), with and without outlier flags, and the product
    library agrees with both at its gates."""
    need_lib()
    z = golden(f"evaluate_{tag}.npz")
    prob = problem_from_golden(z, synth)
    rc = RefCuda(prob)
    try:
        rc.set_level(0)
        c, H, g = rc.evaluate()
        assert abs(c - float(z["cost"])) <= REF_TOL * float(z["cost"])
        assert max_rel(H, z["Hessian"]) <= REF_TOL and max_rel(g, z["gradient"]) <= REF_TOL_G
        c2, _, _ = rc.evaluate(with_h=False)
        assert abs(c2 - float(z["cost_only"])) <= REF_TOL * float(z["cost_only"])
        flags = z["flags"]
        cf, Hf, gf = rc.evaluate(flags=flags, num_bad=int(flags.sum()))
        assert abs(cf - float(z["cost_flagged"])) <= REF_TOL * float(z["cost_flagged"])
        assert max_rel(Hf, z["H_flagged"]) <= REF_TOL and max_rel(gf, z["g_flagged"]) <= REF_TOL_G
    finally:
        rc.close()
    # the C restatement against the same kernels (not only against the CPU harness)
    co, Ho, go, _ = orc.evaluate(prob, 0)
    assert abs(co - c) <= REF_TOL * c and max_rel(Ho, H) <= REF_TOL and max_rel(go, g) <= REF_TOL_G
    # and the product library against the reference's GPU result
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        cm, Hm, gm = ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
    assert abs(cm - c) <= 1e-5 * c and max_rel(Hm, H) <= 1e-5 and max_rel(gm, g) <= 1e-5


def test_reference_cuda_kernels_two_frames_and_flags(pkg, api, O, orc, synth):
    """Two blurred frames on one 3-knot linear spline (each exposure inside its own segment: frame 0 in segment 0, frame 1 in
    segment 1), 512 points, outlier flags: the reference kernels, the oracle and the product library on the same inputs."""
    need_lib()
    prob = synth.make_problem("f2ref", W=192, H=144, levels=1, P0=512, N=8, n_knots=3, k=2, seed=41, margin=20, F=2)
    prob.dt = 1.0
    prob.cap = np.array([0.5, 1.5])
    prob.exp = np.array([0.6, 0.6])
    prob.seg_start = np.array([0, 1], dtype=np.int32)
    flags = np.zeros(prob.levels[0].P, dtype=np.uint8)
    flags[5::17] = 1
    rc = RefCuda(prob)
    try:
        rc.set_level(0)
        for fl, nb in ((None, 0), (flags, int(flags.sum()))):
            c, H, g = rc.evaluate(flags=fl, num_bad=nb)
            co, Ho, go, _ = orc.evaluate(prob, 0, flags=fl, num_bad=nb)
            assert abs(co - c) <= REF_TOL * c and max_rel(Ho, H) <= REF_TOL and max_rel(go, g) <= REF_TOL_G, (abs(co - c) / c, max_rel(Ho, H))
            with pkg.Context(api.limits_for(prob)) as ctx:
                api.upload_problem(ctx, prob)
                if fl is not None:
                    ctx.set_outliers(0, fl, nb)
                cm, Hm, gm = ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
            assert abs(cm - c) <= 1e-5 * c and max_rel(Hm, H) <= 1e-5 and max_rel(gm, g) <= 1e-5
            assert rel(first_step(O, Hm, gm), first_step(O, H, g)) <= 1e-4
    finally:
        rc.close()


def test_reference_cuda_kernels_on_the_bench_configs(pkg, api, O, orc, synth):
    """BASELINE config 1 (2k points, 4 samples) and the coarsest level of config 2: the kernels bench.py times as `gpu_baseline`
    agree with the oracle, so that leg compares like with like."""
    need_lib()
    for name, level in (("C1", 0), ("C2", 3)):
        prob = synth.make_config(name)
        rc = RefCuda(prob)
        try:
            rc.set_level(level)
            c, H, g = rc.evaluate()
        finally:
            rc.close()
        co, Ho, go, _ = orc.evaluate(prob, level)
        assert abs(co - c) <= REF_TOL * c and max_rel(Ho, H) <= REF_TOL and max_rel(go, g) <= REF_TOL_G, name
