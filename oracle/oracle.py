"""TEST INFRASTRUCTURE — ctypes front end of the CPU oracle (oracle/mbavo_oracle.c) and, where it was built,
of oracle/_ref (the reference's own header arithmetic), plus the numpy restatement of the host-side
Levenberg–Marquardt loop of the reference tracker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this module.

Restated host logic (reference file:line):
    solve_normal_equation             src/ba_tracker/solve_normal_equation.h:10-35   (JacobiSVD / LDLT -> numpy)
    computeTrustRegionStep            src/ba_tracker/blur_aware_direct_tracker.cpp:799-831
    Plus_t / Plus_R                   src/core/common/Spline.h:307-330  (Sophus SO3::exp == quaternion exp map)
    optimizePyramidLevel / LM loop    src/ba_tracker/blur_aware_direct_tracker.cpp:590-637, 885-924
    detectOutliersAndUploadToGpu      src/ba_tracker/blur_aware_direct_tracker.cpp:639-699
    constant-velocity prediction      src/ba_tracker/blur_aware_direct_tracker.cpp:120-161, src/core/common/Spline.h:184-219,
                                      Transformation::exp / log = Sophus::SE3d::exp / log (third-party, un-vendored: restated
                                      from the published closed forms, checked against scipy's matrix exponential)
    semi-dense point selection        src/core/feature_detectors/FeatureDetectorSemiDense.cpp:16-59, FeatureDetectorBase.cpp:49-92,
                                      src/core/image_proc/Gradient.h:57-72, blur_aware_direct_tracker.cpp:389-409
    LevenbergMarquardtStrategy        src/ba_tracker/levenberg_marquardt_strategy.cpp:9-44
    TrustRegionStepEvaluator          src/ba_tracker/trust_region_step_evaluator.cpp:45-126
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_ubyte)
_fp = C.POINTER(C.c_float)


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def build(verbose: bool = False) -> None:
    """Compile oracle/libmbavo_oracle.so, and oracle/_ref/* when the reference tree is present."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


class _Lib:
    """Common call surface of libmbavo_oracle.so (prefix mbavo_oracle_) and _ref/libmbavo_ref.so (mbavo_ref_)."""

    def __init__(self, path: str, prefix: str):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self.kind = "reference" if prefix == "mbavo_ref_" else "port"

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def num_threads(self) -> int:
        return int(self.fn("num_threads")())

    def set_num_threads(self, n: int) -> None:
        self.fn("set_num_threads")(C.c_int(n))

    # -- stage 1
    def virtual_poses(self, N, cap, exp, k, t0, dt, knots_t, knots_R, jac=True):
        cap = np.ascontiguousarray(cap, dtype=np.float64)
        exp = np.ascontiguousarray(exp, dtype=np.float64)
        kt = np.ascontiguousarray(knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64)
        F = cap.shape[0]
        poses = np.zeros((F, N, 7))
        Jt = np.zeros((F, N, 3, 3 * k)) if jac else None
        JR = np.zeros((F, N, 4, 3 * k)) if jac else None
        seg = np.zeros((F, N), dtype=np.int32)
        f = self.fn("virtual_poses")
        f.restype = C.c_int
        if self.prefix == "mbavo_ref_":
            rc = f(C.c_int(N), C.c_int(F), _ptr(cap, _dp), _ptr(exp, _dp), C.c_int(k), C.c_double(t0), C.c_double(dt),
                   _ptr(kt, _dp), _ptr(kR, _dp), _ptr(poses, _dp), _ptr(Jt, _dp), _ptr(JR, _dp), _ptr(seg, _ip))
        else:
            rc = f(C.c_int(N), C.c_int(F), _ptr(cap, _dp), _ptr(exp, _dp), C.c_int(k), C.c_double(t0), C.c_double(dt),
                   _ptr(kt, _dp), _ptr(kR, _dp), C.c_int(kt.shape[0]), _ptr(poses, _dp), _ptr(seg, _ip),
                   _ptr(Jt, _dp), _ptr(JR, _dp))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}virtual_poses rc={rc}")
        return poses, seg, Jt, JR

    # -- stage 2
    def local_patches(self, N, poses, xy, z, fx, fy, cx, cy):
        F = poses.shape[0]
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        P = xy.shape[0]
        out = np.zeros((F, P, 2))
        f = self.fn("local_patches")
        f.restype = C.c_int
        f(C.c_int(N), C.c_int(F), _ptr(np.ascontiguousarray(poses), _dp), _ptr(xy, _dp), _ptr(z, _dp), C.c_int(P),
          C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), _ptr(out, _dp))
        return out

    # -- one sample: compute_pixel_intensity<double>
    def pixel_intensity(self, I, dIxy, pose, D, fx, fy, cx, cy, X, Y, want_J=True):
        H, W = I.shape
        inten = C.c_double(0)
        J = np.zeros(7) if want_J else None
        f = self.fn("pixel_intensity")
        f.restype = C.c_int
        ok = f(_ptr(I, _u8p), _ptr(dIxy, _fp), C.c_int(H), C.c_int(W), _ptr(np.ascontiguousarray(pose, dtype=np.float64), _dp),
               C.c_double(D), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), C.c_double(X),
               C.c_double(Y), C.byref(inten), _ptr(J, _dp))
        return bool(ok), inten.value, J

    # -- whole evaluation
    def evaluate(self, prob, level: int, knots_t=None, knots_R=None, flags=None, num_bad: int = 0,
                 with_hessian: bool = True, want_patch_costs: bool = True, centres_in=None, centres_out=None):
        """-> (cost, H, g, patch_costs).  centres_in / centres_out ([F, P, 2], oracle port only) hold the patch centres
        fixed / report them (finite-difference checks)."""
        lv = prob.levels[level]
        kt = np.ascontiguousarray(prob.knots_t if knots_t is None else knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(prob.knots_R if knots_R is None else knots_R, dtype=np.float64)
        n = kt.shape[0]
        F = prob.F
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in lv.cur_I])
        cost = C.c_double(0)
        H = np.zeros((6 * n, 6 * n)) if with_hessian else None
        g = np.zeros(6 * n) if with_hessian else None
        pc = np.zeros((F, lv.P)) if want_patch_costs else None
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        seg = np.ascontiguousarray(prob.seg_start, dtype=np.int32)
        f = self.fn("evaluate")
        f.restype = C.c_int
        common = [C.c_int(lv.N), C.c_int(F), _ptr(lv.ref_I, _u8p), _ptr(lv.ref_dIxy, _fp), cur,
                  _ptr(np.ascontiguousarray(prob.cap), _dp), _ptr(np.ascontiguousarray(prob.exp), _dp),
                  _ptr(lv.xy, _dp), _ptr(lv.z, _dp), C.c_int(lv.P), _ptr(lv.pattern, _ip), C.c_int(lv.S),
                  _ptr(fl, _u8p), C.c_int(num_bad), C.c_double(lv.fx), C.c_double(lv.fy), C.c_double(lv.cx),
                  C.c_double(lv.cy), C.c_int(lv.H), C.c_int(lv.W), C.c_int(prob.k), C.c_double(prob.t0),
                  C.c_double(prob.dt), _ptr(kt, _dp), _ptr(kR, _dp)]
        if self.prefix == "mbavo_ref_":
            rc = f(*common, _ptr(seg, _ip), C.c_int(n), C.c_double(prob.huber_a), C.byref(cost), _ptr(H, _dp),
                   _ptr(g, _dp), _ptr(pc, _dp))
        elif centres_in is not None or centres_out is not None:
            f = self.fn("evaluate_fixed_centres")
            f.restype = C.c_int
            ci = None if centres_in is None else np.ascontiguousarray(centres_in, dtype=np.float64)
            rc = f(*common, C.c_int(n), C.c_double(prob.huber_a), C.byref(cost), _ptr(H, _dp), _ptr(g, _dp),
                   _ptr(pc, _dp), _ptr(ci, _dp), _ptr(centres_out, _dp))
        else:
            rc = f(*common, C.c_int(n), C.c_double(prob.huber_a), C.byref(cost), _ptr(H, _dp), _ptr(g, _dp),
                   _ptr(pc, _dp))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}evaluate rc={rc}")
        return cost.value, H, g, pc


class OracleLib(_Lib):
    def __init__(self):
        path = os.path.join(_HERE, "libmbavo_oracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path, "mbavo_oracle_")

    def pixel_residuals(self, prob, level: int, knots_t=None, knots_R=None, with_jacobian=True):
        lv = prob.levels[level]
        kt = np.ascontiguousarray(prob.knots_t if knots_t is None else knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(prob.knots_R if knots_R is None else knots_R, dtype=np.float64)
        n = kt.shape[0]
        F = prob.F
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in lv.cur_I])
        r = np.zeros((F, lv.P, lv.S))
        cap_cols = 6 * n
        J = np.zeros((F, lv.P, lv.S, cap_cols)) if with_jacobian else None
        kmin, NK = C.c_int(0), C.c_int(0)
        f = self.fn("pixel_residuals")
        f.restype = C.c_int
        # the C side writes rows of width 6*NK contiguously; allocate for the worst case and reshape afterwards
        rc = f(C.c_int(lv.N), C.c_int(F), _ptr(lv.ref_I, _u8p), _ptr(lv.ref_dIxy, _fp), cur,
               _ptr(np.ascontiguousarray(prob.cap), _dp), _ptr(np.ascontiguousarray(prob.exp), _dp), _ptr(lv.xy, _dp),
               _ptr(lv.z, _dp), C.c_int(lv.P), _ptr(lv.pattern, _ip), C.c_int(lv.S), C.c_double(lv.fx),
               C.c_double(lv.fy), C.c_double(lv.cx), C.c_double(lv.cy), C.c_int(lv.H), C.c_int(lv.W), C.c_int(prob.k),
               C.c_double(prob.t0), C.c_double(prob.dt), _ptr(kt, _dp), _ptr(kR, _dp), C.c_int(n), C.byref(kmin),
               C.byref(NK), _ptr(r, _dp), _ptr(J, _dp), C.c_int(cap_cols))
        if rc != 0:
            raise RuntimeError(f"mbavo_oracle_pixel_residuals rc={rc}")
        if J is not None:
            d = 6 * NK.value
            J = J.reshape(-1)[: F * lv.P * lv.S * d].reshape(F, lv.P, lv.S, d)
        return r, J, kmin.value, NK.value

    def image_gradient(self, I):
        H, W = I.shape
        out = np.zeros((H, W, 2), dtype=np.float32)
        self.fn("image_gradient")(_ptr(np.ascontiguousarray(I), _u8p), C.c_int(H), C.c_int(W), _ptr(out, _fp))
        return out

    def pyramid_down(self, I):
        H, W = I.shape
        out = np.zeros((H // 2, W // 2), dtype=np.uint8)
        self.fn("pyramid_down")(_ptr(np.ascontiguousarray(I), _u8p), C.c_int(H), C.c_int(W), _ptr(out, _u8p))
        return out

    def warp_mean(self, I, D, fx, fy, cx, cy, poses_tq):
        """Mean of the plane-induced warps of I through the given poses ((n, 7): tx ty tz qx qy qz qw), oracle port only."""
        I = np.ascontiguousarray(I, dtype=np.uint8)
        poses = np.ascontiguousarray(poses_tq, dtype=np.float64).reshape(-1, 7)
        H, W = I.shape
        out = np.zeros((H, W), dtype=np.uint8)
        f = self.fn("warp_mean")
        f.restype = C.c_int
        rc = f(_ptr(I, _u8p), C.c_int(H), C.c_int(W), C.c_double(D), C.c_double(fx), C.c_double(fy), C.c_double(cx),
               C.c_double(cy), _ptr(poses, _dp), C.c_int(poses.shape[0]), _ptr(out, _u8p))
        if rc != 0:
            raise RuntimeError(f"warp_mean rc={rc}")
        return out

    def synthesize_blurred(self, I, D, fx, fy, cx, cy, k, t0, dt, knots_t, knots_R, cap, exp, num_samples):
        H, W = I.shape
        kt = np.ascontiguousarray(knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64)
        out = np.zeros((H, W), dtype=np.uint8)
        f = self.fn("synthesize_blurred")
        f.restype = C.c_int
        rc = f(_ptr(np.ascontiguousarray(I), _u8p), C.c_int(H), C.c_int(W), C.c_double(D), C.c_double(fx),
               C.c_double(fy), C.c_double(cx), C.c_double(cy), C.c_int(k), C.c_double(t0), C.c_double(dt),
               _ptr(kt, _dp), _ptr(kR, _dp), C.c_int(kt.shape[0]), C.c_double(cap), C.c_double(exp),
               C.c_int(num_samples), _ptr(out, _u8p))
        if rc != 0:
            raise RuntimeError(f"synthesize_blurred rc={rc}")
        return out


class RefLib(_Lib):
    """oracle/_ref/libmbavo_ref.so — present only where it was built from /root/reference (or shipped prebuilt)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libmbavo_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        super().__init__(path, "mbavo_ref_")

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libmbavo_ref.so"))

    def spline_pose(self, k, t0, dt, knots_t, knots_R, t):
        """One pose by the reference's spline functors -> [tx ty tz qx qy qz qw]."""
        kt = np.ascontiguousarray(knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64)
        out = np.zeros(7)
        f = self.fn("spline_pose")
        f.restype = C.c_int
        if f(C.c_int(k), C.c_double(t0), C.c_double(dt), _ptr(kt, _dp), _ptr(kR, _dp), C.c_double(t), _ptr(out, _dp)) != 0:
            raise RuntimeError("mbavo_ref_spline_pose failed")
        return out

    def synthesize_blurred(self, I, D, fx, fy, cx, cy, k, t0, dt, knots_t, knots_R, cap, exp, num_samples):
        H, W = I.shape
        kt = np.ascontiguousarray(knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64)
        out = np.zeros((H, W), dtype=np.uint8)
        f = self.fn("synthesize_blurred")
        f.restype = C.c_int
        f(_ptr(np.ascontiguousarray(I), _u8p), C.c_int(H), C.c_int(W), C.c_double(D), C.c_double(fx), C.c_double(fy),
          C.c_double(cx), C.c_double(cy), C.c_int(k), C.c_double(t0), C.c_double(dt), _ptr(kt, _dp), _ptr(kR, _dp),
          C.c_double(cap), C.c_double(exp), C.c_int(num_samples), _ptr(out, _u8p))
        return out


def keyframe_stats(xy, z, fx, fy, cx, cy, poses_tq):
    """isKeyframe statistics (blur_aware_direct_tracker.cpp:205-248), numpy fp64 restatement: -> (avg_flow, avg_kernel_len).
    poses_tq: (3, 7) = T_cur2ref at the capture time, at -half and at +half the exposure (tx ty tz qx qy qz qw)."""
    xy = np.asarray(xy, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    Pr = np.stack([z * ((xy[:, 0] - cx) / fx), z * ((xy[:, 1] - cy) / fy), z], axis=1)  # CameraPinhole::unproject
    uv = []
    for p in np.asarray(poses_tq, dtype=np.float64).reshape(3, 7):
        t, (qx, qy, qz, qw) = p[:3], p[3:]
        R = np.array([[qw * qw + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)],
                      [2 * (qx * qy + qw * qz), qw * qw - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qw * qx)],
                      [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz]])
        Pc = (Pr - t) @ R  # T.inverse() * P = R^T (P - t)
        uv.append(np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], axis=1))
    flow = ((uv[0] - xy) ** 2).sum()
    kern = ((uv[1] - uv[2]) ** 2).sum()
    n = xy.shape[0]
    return float(np.sqrt(np.float32(flow / n))), float(np.sqrt(np.float32(kern / n)))


# ---------------------------------------------------------------------------------------------------------
# per-frame trajectory bookkeeping (numpy, rotation-matrix formulation — independent of the product's quaternion code)
# ---------------------------------------------------------------------------------------------------------

def q_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def se3_exp(tangent):
    """Sophus::SE3d::exp, tangent = [upsilon, omega] -> (t, q (x, y, z, w))."""
    ups, om = np.asarray(tangent[:3], np.float64), np.asarray(tangent[3:], np.float64)
    th = np.linalg.norm(om)
    q = so3_exp_quat(om)
    Om = hat(om)
    if th < 1e-10:
        V = q_to_R(q)
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * (Om @ Om)
    return V @ ups, q


def se3_log(t, q):
    """Sophus::SE3d::log -> [upsilon, omega]."""
    v, w = np.asarray(q[:3], np.float64), q[3]
    n = np.linalg.norm(v)
    if n < 1e-10:
        f = 2.0 / w - 2.0 * n * n / w ** 3
    elif abs(w) < 1e-10:
        f = (np.pi if w > 0 else -np.pi) / n
    else:
        f = 2.0 * np.arctan(n / w) / n
    om = f * v
    th = f * n
    Om = hat(om)
    if abs(th) < 1e-10:
        Vi = np.eye(3) - 0.5 * Om + (Om @ Om) / 12.0
    else:
        Vi = np.eye(3) - 0.5 * Om + (1 - th * np.cos(th / 2) / (2 * np.sin(th / 2))) / th ** 2 * (Om @ Om)
    return np.concatenate([Vi @ np.asarray(t, np.float64), om])


def transform_by_right(knots_t, knots_R, dq, dt):
    """SplineSE3::TransformByRight (Spline.h:212-219)."""
    kt = np.array([q_to_R(q / np.linalg.norm(q)) @ dt * 1.0 + t for q, t in zip(knots_R, knots_t)])
    kR = np.array([q_mul(q, dq) for q in knots_R])
    return kt, kR


def predict_spline(knots_t, knots_R, velocity, dt_frame):
    """blur_aware_direct_tracker.cpp:120-145."""
    t, q = se3_exp(np.asarray(velocity, np.float64) * dt_frame)
    return transform_by_right(knots_t, knots_R, q, t)


def frame_velocity(prev_t, prev_q, cur_t, cur_q, dt_frame):
    """blur_aware_direct_tracker.cpp:155-161: log(T_prev^-1 T_cur) / dt."""
    Rp = q_to_R(prev_q)
    qi = np.array([-prev_q[0], -prev_q[1], -prev_q[2], prev_q[3]])
    return se3_log(Rp.T @ (np.asarray(cur_t) - np.asarray(prev_t)), q_mul(qi, cur_q)) / dt_frame


# ---------------------------------------------------------------------------------------------------------
# semi-dense host-map point selection (numpy restatement; pinned against oracle/_ref/libmbavo_refselect.so)
# ---------------------------------------------------------------------------------------------------------

def pyramid_numpy(I0, n_levels):
    """ImagePyramid<uchar>::computePyramid (ImagePyramid.h:59-99): level l has H0 / 2^l x W0 / 2^l pixels, each the truncated
    quarter of the float sum of its 2 x 2 parents."""
    out = [np.ascontiguousarray(I0, dtype=np.uint8)]
    H0, W0 = out[0].shape
    for lv in range(1, n_levels):
        p = out[-1].astype(np.uint32)
        H, W = H0 // (1 << lv), W0 // (1 << lv)
        s = p[0:2 * H:2, 0:2 * W:2] + p[0:2 * H:2, 1:2 * W:2] + p[1:2 * H:2, 0:2 * W:2] + p[1:2 * H:2, 1:2 * W:2]
        out.append((s >> 2).astype(np.uint8))  # T(0.25f * sum): the sum (<= 1020) is exact in float
    return out


def gradient_magnitude(I):
    """Magnitude image of compute_image_gradients (Gradient.h:33-72) for one channel: sqrt(dx^2 + dy^2) of the halved central
    differences (float products and sum — exact for 8-bit inputs — square root in double, stored as float), 0 on the border."""
    f = I.astype(np.float32)
    H, W = f.shape
    dx = np.zeros((H, W), np.float32)
    dy = np.zeros((H, W), np.float32)
    dx[1:-1, 1:-1] = np.float32(0.5) * (f[1:-1, 2:] - f[1:-1, :-2])
    dy[1:-1, 1:-1] = np.float32(0.5) * (f[2:, 1:-1] - f[:-2, 1:-1])
    return np.sqrt((dx * dx + dy * dy).astype(np.float64)).astype(np.float32)


def select_points(I0, n_levels, score_threshold, cell_H, cell_W, depth_z):
    """FeatureDetectorSemiDense::detect + FeatureDetectorBase::gridSelection + the depth look-up of tmpProcessKeyframe.
    -> per level (xy (P, 2) float64 in level coordinates, z (P,) float64), points in cell order."""
    H0, W0 = I0.shape
    depth_z = np.ascontiguousarray(depth_z, dtype=np.float32)
    thr = np.float32(score_threshold)
    out = []
    for lv, I in enumerate(pyramid_numpy(I0, n_levels)):
        mag = gradient_magnitude(I)
        H_lv, W_lv = H0 // int(2.0 ** lv), W0 // int(2.0 ** lv)        # FeatureDetectorBase.cpp:57-59
        ch, cw = int(cell_H / 1.414 ** lv), int(cell_W / 1.414 ** lv)  # :60-61
        ncw = W_lv // cw + 1                                           # :63-64
        ys, xs = np.nonzero(mag > thr)                                 # row-major scan, SemiDense.cpp:29-43
        resp = mag[ys, xs]
        # float division then truncation, as `pt.y / cell_H_lv` (:70-72)
        cell = (ys.astype(np.float32) / np.float32(ch)).astype(np.int64) * ncw + (xs.astype(np.float32) / np.float32(cw)).astype(np.int64)
        order = np.lexsort((np.arange(len(cell)), -resp.astype(np.float64), cell))  # per cell: strongest, first in scan order on ties
        first = np.ones(len(order), bool)
        first[1:] = cell[order][1:] != cell[order][:-1]
        pick = order[first]
        pick = pick[resp[pick] >= np.float32(1e-6)]                    # :85-88
        px, py = xs[pick].astype(np.float32), ys[pick].astype(np.float32)
        scale = 2.0 ** lv
        x0 = (px.astype(np.float64) * scale + 0.5).astype(np.int64)    # tracker.cpp:397-398
        y0 = (py.astype(np.float64) * scale + 0.5).astype(np.int64)
        z = depth_z[y0, x0]
        keep = ~(z.astype(np.float64) < 1e-2)                          # :401-404
        out.append((np.stack([px[keep], py[keep]], axis=1).astype(np.float64), z[keep].astype(np.float64)))
    return out


class RefSelect:
    """oracle/_ref/libmbavo_refselect.so — the reference's own detector sources, where they were built."""

    PATH = os.path.join(_HERE, "_ref", "libmbavo_refselect.so")

    @staticmethod
    def available() -> bool:
        return os.path.exists(RefSelect.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)

    def select_points(self, I0, n_levels, score_threshold, cell_H, cell_W, depth_z, max_points=None, want_mag=False):
        H0, W0 = I0.shape
        max_points = max_points or H0 * W0
        xy = np.zeros((n_levels, max_points, 2), np.float64)
        z = np.zeros((n_levels, max_points), np.float64)
        cnt = np.zeros(n_levels, np.int32)
        mag = np.zeros((H0, W0), np.float32) if want_mag else None
        self.lib.mbavo_refselect_points(_ptr(np.ascontiguousarray(I0, dtype=np.uint8), _u8p), C.c_int(H0), C.c_int(W0), C.c_int(n_levels),
                                        C.c_float(score_threshold), C.c_int(cell_H), C.c_int(cell_W),
                                        _ptr(np.ascontiguousarray(depth_z, dtype=np.float32), _fp), C.c_int(max_points),
                                        _ptr(xy, _dp), _ptr(z, _dp), _ptr(cnt, _ip), _ptr(mag, _fp))
        res = [(xy[l, :cnt[l]].copy(), z[l, :cnt[l]].copy()) for l in range(n_levels)]
        return (res, mag) if want_mag else res

    def pyramid(self, I0, n_levels):
        """The reference's own computePyramid + compute_image_gradients -> [(image u8 (H, W), gradient float32 (H, W, 2))] per level."""
        H0, W0 = I0.shape
        sizes = [(H0 // (1 << l), W0 // (1 << l)) for l in range(n_levels)]
        total = sum(h * w for h, w in sizes)
        levels = np.zeros(total, np.uint8)
        grads = np.zeros(2 * total, np.float32)
        self.lib.mbavo_refselect_pyramid(_ptr(np.ascontiguousarray(I0, dtype=np.uint8), _u8p), C.c_int(H0), C.c_int(W0), C.c_int(n_levels),
                                         _ptr(levels, _u8p), _ptr(grads, _fp))
        out, off = [], 0
        for h, w in sizes:
            out.append((levels[off:off + h * w].reshape(h, w).copy(), grads[2 * off:2 * (off + h * w)].reshape(h, w, 2).copy()))
            off += h * w
        return out


def best_cpu_lib() -> _Lib:
    """oracle/_ref when present (kind 'reference'), else the C port (kind 'port')."""
    return RefLib() if RefLib.available() else OracleLib()


# ---------------------------------------------------------------------------------------------------------
# host-side solver logic (numpy)
# ---------------------------------------------------------------------------------------------------------
def q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def so3_exp_quat(w):
    """Sophus::SO3d::exp(w).unit_quaternion() (Spline.h:302, 326): half-angle exponential map, Taylor branch for
    tiny angles (Sophus uses theta^2 < eps^2 with the same series)."""
    w = np.asarray(w, dtype=np.float64)
    t2 = float(w @ w)
    if t2 < 1e-20:
        t4 = t2 * t2
        return np.concatenate([(0.5 - t2 / 48.0 + t4 / 3840.0) * w, [1.0 - t2 / 8.0 + t4 / 384.0]])
    t = np.sqrt(t2)
    return np.concatenate([np.sin(0.5 * t) / t * w, [np.cos(0.5 * t)]])


def plus(knots_t, knots_R, delta):
    """Spline.h:307-330: candidate = knots [+] delta, delta = [dt_0..dt_{n-1}, dw_0..dw_{n-1}]; no re-normalisation."""
    n = knots_t.shape[0]
    ct = knots_t + delta[: 3 * n].reshape(n, 3)
    cR = np.stack([q_mul(knots_R[i], so3_exp_quat(delta[3 * n + 3 * i: 3 * n + 3 * i + 3])) for i in range(n)])
    return ct, cR


def solve_normal_equation(A, b, solver_type: str = "SVD_JACOBI"):
    """solve_normal_equation.h:16-34: x = -A^-1 b by SVD (type 0) or LDL^T (type 1)."""
    if solver_type == "SVD_JACOBI":
        U, s, Vt = np.linalg.svd(A)
        # Eigen's JacobiSVD::solve drops singular values below eps * max(rows, cols) * s_max
        tol = np.finfo(np.float64).eps * max(A.shape) * s[0]
        inv = np.where(s > tol, 1.0 / np.where(s > tol, s, 1.0), 0.0)
        x = Vt.T @ (inv * (U.T @ b))
    elif solver_type == "LDLT":
        x = np.linalg.solve(A, b)
    else:
        raise ValueError(solver_type)
    return -x


def trust_region_step(H, g, radius, solver_type="SVD_JACOBI"):
    """tracker.cpp:799-831.  H is damped IN PLACE (the damping compounds over rejected steps)."""
    d = np.diag_indices_from(H)
    H[d] += H[d] * (1.0 / radius)
    step = solve_normal_equation(H, g, solver_type)
    model = -(g @ step + 0.5 * step @ H @ step)
    return step, model


class LMStrategy:
    """levenberg_marquardt_strategy.cpp:9-44"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.radius, self.min_radius, self.max_radius, self.decrease = 1e4, 10.0, 1e32, 2.0

    def accepted(self, q):
        self.radius = self.radius / max(1.0 / 3.0, 1.0 - (2.0 * q - 1.0) ** 3)
        self.radius = max(min(self.max_radius, self.radius), self.min_radius)
        self.decrease = 2.0

    def rejected(self):
        self.radius = self.radius / self.decrease
        self.radius = max(min(self.max_radius, self.radius), self.min_radius)
        self.decrease *= 2.0


class StepEvaluator:
    """trust_region_step_evaluator.cpp:45-126 (Ceres' non-monotonic step evaluator)."""

    def __init__(self, max_nonmonotonic=5):
        self.max_nonmonotonic = max_nonmonotonic

    def reset(self, c):
        self.minimum = self.current = self.reference = self.candidate = c
        self.acc_ref = self.acc_cand = 0.0
        self.n_nonmono = 0

    def quality(self, cost, model):
        if cost >= np.finfo(np.float64).max:
            return np.finfo(np.float64).min
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = (self.current - cost) / model
            hist = (self.reference - cost) / (self.acc_ref + model)
        return max(rel, hist)

    def accepted(self, cost, model):
        self.current = cost
        self.acc_cand += model
        self.acc_ref += model
        if self.current < self.minimum:
            self.minimum = self.current
            self.n_nonmono = 0
            self.candidate = self.current
            self.acc_cand = 0.0
        else:
            self.n_nonmono += 1
            if self.current > self.candidate:
                self.candidate = self.current
                self.acc_cand = 0.0
        if self.n_nonmono == self.max_nonmonotonic:
            self.reference = self.candidate
            self.acc_ref = self.acc_cand


def detect_outliers(patch_costs, flags, k_sigma):
    """tracker.cpp:639-699: mean / variance over patches with cost >= 1e-8, flag every patch (zero-cost ones included)
    with |c - mu| > k_sigma * sqrtf(var).  Flags are sticky; returns the number flagged in THIS call."""
    c = np.asarray(patch_costs, dtype=np.float64).reshape(-1)
    sel = c[c >= 1e-8]
    mu = sel.sum() / sel.size
    var = ((sel - mu) ** 2).sum() / sel.size
    bad = np.abs(c - mu) > k_sigma * float(np.sqrt(np.float32(var)))
    flags[bad] = 1
    return int(bad.sum())


@dataclass
class LMTrace:
    costs: List[float] = field(default_factory=list)        # evaluation-point cost after every accepted step
    decisions: List[str] = field(default_factory=list)      # 'A' accepted, 'R' rejected, 'I' invalid step
    qualities: List[float] = field(default_factory=list)
    first_step: Optional[np.ndarray] = None
    num_iterations: int = 0
    num_bad: int = 0


def optimize_level(lib: _Lib, prob, level: int, knots_t, knots_R, max_num_iterations=50, min_step_quality=0.5,
                   min_abs_cost_decrease=1e-3, solver_type="SVD_JACOBI", max_nonmonotonic=5):
    """optimizePyramidLevel (tracker.cpp:590-637) with evaluate() supplied by `lib`."""
    lv = prob.levels[level]
    kt, kR = knots_t.copy(), knots_R.copy()
    flags = np.zeros(lv.P, dtype=np.uint8)
    num_bad = 0
    trace = LMTrace()
    eval_cost, H, g, _ = lib.evaluate(prob, level, kt, kR, flags, num_bad, with_hessian=True)
    lm, ev = LMStrategy(), StepEvaluator(max_nonmonotonic)
    ev.reset(eval_cost)
    trace.costs.append(eval_cost)
    it, abs_decrease = 0, 1e10
    while True:
        it += 1                                     # finalizeIterationAndCheckIfMinimizerCanContinue, :910-924
        if it > max_num_iterations or abs_decrease < min_abs_cost_decrease:
            break
        step, model = trust_region_step(H, g, lm.radius, solver_type)
        if trace.first_step is None:
            trace.first_step = step.copy()
        if model < 0:
            lm.rejected()
            trace.decisions.append("I")
            continue
        ct, cR = plus(kt, kR, step)
        cand_cost, _, _, pc = lib.evaluate(prob, level, ct, cR, flags, num_bad, with_hessian=False)
        abs_decrease = eval_cost - cand_cost
        q = ev.quality(cand_cost, model)
        trace.qualities.append(q)
        if q > min_step_quality and cand_cost < eval_cost:
            num_bad = detect_outliers(pc, flags, prob.max_chi_square_error)
            kt, kR = ct, cR
            eval_cost, H, g, _ = lib.evaluate(prob, level, kt, kR, flags, num_bad, with_hessian=True)
            lm.accepted(q)
            ev.accepted(eval_cost, model)
            trace.decisions.append("A")
            trace.costs.append(eval_cost)
        else:
            lm.rejected()
            trace.decisions.append("R")
    trace.num_iterations = it - 1
    trace.num_bad = num_bad
    return kt, kR, trace


def optimize_trajectory(lib: _Lib, prob, **kw):
    """optimizeTrajectory (tracker.cpp:544-588): coarse -> fine over the pyramid."""
    kt, kR = prob.knots_t.copy(), prob.knots_R.copy()
    traces = []
    for level in reversed(range(len(prob.levels))):
        kt, kR, tr = optimize_level(lib, prob, level, kt, kR, **kw)
        traces.append(tr)
    return kt, kR, traces
