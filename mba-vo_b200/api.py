"""ctypes binding of include/mbavo.h — one Python method per C entry point, no logic of its own."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MAX_LEVELS = 8
MAX_FRAMES = 16
MEM_HOST, MEM_DEVICE = 0, 1
SOLVER_SVD_JACOBI, SOLVER_LDLT = 0, 1

# every symbol include/mbavo.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "mbavo_last_error", "mbavo_version", "mbavo_create", "mbavo_destroy", "mbavo_set_stream", "mbavo_set_frame_times",
    "mbavo_set_level", "mbavo_set_keyframe_pyramid", "mbavo_set_live_pyramid", "mbavo_set_level_points", "mbavo_set_points_pyramid",
    "mbavo_set_frame", "mbavo_debug_dump", "mbavo_debug_cost_tma",
    "mbavo_set_live_images", "mbavo_set_outliers", "mbavo_set_num_bad", "mbavo_evaluate", "mbavo_patch_costs",
    "mbavo_detect_outliers", "mbavo_packed_len", "mbavo_evaluate_async", "mbavo_unpack", "mbavo_trust_region_step",
    "mbavo_spline_plus", "mbavo_gn_iteration", "mbavo_gn_sweep", "mbavo_lm_default_options", "mbavo_optimize_level", "mbavo_kernel_launches",
    "mbavo_device_sweeps", "mbavo_persistent_sweeps", "mbavo_sweep_pass_times", "mbavo_enable_kernel_timing", "mbavo_last_kernel_ms", "mbavo_level_uses_texels", "mbavo_shard_export",
    "mbavo_shard_connect", "mbavo_shard_disconnect", "mbavo_shard_set_global_points", "mbavo_synthesize_blurred",
    "mbavo_keyframe_stats", "mbavo_select_points", "mbavo_get_points", "mbavo_se3_exp", "mbavo_se3_log",
    "mbavo_spline_pose", "mbavo_spline_transform_by_right", "mbavo_spline_transform_to", "mbavo_predict_spline", "mbavo_frame_velocity",
    "mbavo_tracker_init", "mbavo_track_frame", "mbavo_tracker_new_keyframe", "mbavo_is_keyframe",
]
IPC_HANDLE_BYTES = 64


class MbavoError(RuntimeError):
    pass


class _PointSelection(C.Structure):
    _fields_ = [("score_threshold", C.c_float), ("cell_H", C.c_int), ("cell_W", C.c_int), ("depth_mem", C.c_int),
                ("depth_z", C.c_void_p), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("pattern_xy", C.c_void_p), ("patch_size", C.c_int), ("num_virtual_poses", C.c_int)]


class _Limits(C.Structure):
    _fields_ = [("device", C.c_int), ("max_num_frames", C.c_int), ("max_num_virtual_poses_per_frame", C.c_int),
                ("max_num_keypoints", C.c_int), ("max_patch_size", C.c_int), ("max_num_ctrl_knots", C.c_int)]


class _Level(C.Structure):
    _fields_ = [("mem", C.c_int), ("H", C.c_int), ("W", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("ref_I", C.c_void_p), ("ref_dIxy", C.c_void_p),
                ("cur_I", C.POINTER(C.c_void_p)), ("n_frames", C.c_int), ("keypoint_xy", C.c_void_p),
                ("keypoint_xy_stride", C.c_int), ("keypoint_xy_offset", C.c_int), ("keypoint_z", C.c_void_p),
                ("num_keypoints", C.c_int), ("pattern_xy", C.c_void_p), ("patch_size", C.c_int),
                ("num_virtual_poses", C.c_int), ("ext_outlier_flags", C.c_void_p), ("ext_patch_cost", C.c_void_p),
                ("ext_patch_cost_stride", C.c_int)]


class _LevelPoints(C.Structure):
    _fields_ = [("mem", C.c_int), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("keypoint_xy", C.c_void_p), ("keypoint_xy_stride", C.c_int), ("keypoint_xy_offset", C.c_int),
                ("keypoint_z", C.c_void_p), ("num_keypoints", C.c_int), ("pattern_xy", C.c_void_p), ("patch_size", C.c_int),
                ("num_virtual_poses", C.c_int)]


class _Spline(C.Structure):
    _fields_ = [("spline_deg_k", C.c_int), ("start_time", C.c_double), ("sample_dt", C.c_double),
                ("num_ctrl_knots", C.c_int), ("knots_t", C.POINTER(C.c_double)), ("knots_R", C.POINTER(C.c_double))]


class _DebugOut(C.Structure):
    _fields_ = [("poses_tq", C.c_void_p), ("blend_weights", C.c_void_p), ("theta", C.c_void_p), ("segment_start_knot", C.c_void_p),
                ("patch_centres", C.c_void_p), ("residuals", C.c_void_p), ("jacobians", C.c_void_p), ("kmin", C.c_int),
                ("knot_window", C.c_int), ("spline_deg_k", C.c_int), ("cost", C.c_double)]


class _LmOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int), ("min_step_quality", C.c_double), ("min_abs_cost_decrease", C.c_double),
                ("solver_type", C.c_int), ("max_consecutive_nonmonotonic_steps", C.c_int),
                ("max_chi_square_error", C.c_double), ("huber_a", C.c_double)]


class _LmSummary(C.Structure):
    _fields_ = [("num_iterations", C.c_int), ("num_accepted", C.c_int), ("num_rejected", C.c_int),
                ("num_invalid", C.c_int), ("num_evaluations", C.c_int), ("num_bad_keypoints", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("first_step", C.c_double * 96),
                ("decisions", C.c_char * 64)]


class _Tracker(C.Structure):
    _fields_ = [("spline_deg_k", C.c_int), ("num_ctrl_knots", C.c_int), ("sample_dt", C.c_double), ("start_time", C.c_double),
                ("knots_t", C.c_double * 48), ("knots_R", C.c_double * 64), ("velocity", C.c_double * 6),
                ("prev_t", C.c_double * 3), ("prev_q", C.c_double * 4), ("prev_timestamp", C.c_double),
                ("keyframe_t", C.c_double * 3), ("keyframe_q", C.c_double * 4)]


class _FrameResult(C.Structure):
    _fields_ = [("t_cur2key", C.c_double * 3), ("q_cur2key", C.c_double * 4), ("t_cur2world", C.c_double * 3),
                ("q_cur2world", C.c_double * 4), ("avg_flow", C.c_double), ("avg_kernel_len", C.c_double),
                ("levels_run", C.c_int), ("levels", _LmSummary * 8)]


@dataclass
class Limits:
    max_num_frames: int = 1
    max_num_virtual_poses_per_frame: int = 64
    max_num_keypoints: int = 500
    max_patch_size: int = 128
    max_num_ctrl_knots: int = 16
    device: int = -1


def library_path() -> str:
    """lib/libmbavo_b200.so; MBAVO_LIBRARY selects another build of the same sources (kernel-tuning experiments)."""
    return os.environ.get("MBAVO_LIBRARY") or os.path.join(_HERE, "lib", "libmbavo_b200.so")


_LIB = None


def load_library() -> C.CDLL:
    """Load lib/libmbavo_b200.so.  Raises if it has not been built — there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise MbavoError(f"{path} is missing: build it with `make -C mba-vo_b200` or __graft_entry__.build()")
    lib = C.CDLL(path)
    lib.mbavo_last_error.restype = C.c_char_p
    lib.mbavo_kernel_launches.restype = C.c_longlong
    lib.mbavo_device_sweeps.restype = C.c_longlong
    lib.mbavo_persistent_sweeps.restype = C.c_longlong
    lib.mbavo_last_kernel_ms.restype = C.c_float
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)
    _LIB = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One mbavo_ctx: one logical tracker on one GPU."""

    def __init__(self, limits: Limits):
        self.lib = load_library()
        self._h = C.c_void_p()
        lim = _Limits(limits.device, limits.max_num_frames, limits.max_num_virtual_poses_per_frame,
                      limits.max_num_keypoints, limits.max_patch_size, limits.max_num_ctrl_knots)
        self._check(self.lib.mbavo_create(C.byref(lim), C.byref(self._h)))
        self.limits = limits
        self._keep = {}

    def _check(self, rc: int):
        if rc != 0:
            raise MbavoError(f"mbavo error {rc}: {self.lib.mbavo_last_error().decode()}")

    def close(self):
        if self._h:
            self.lib.mbavo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs -------------------------------------------------------------------------------------------
    def set_stream(self, stream_ptr: Optional[int]):
        self._check(self.lib.mbavo_set_stream(self._h, C.c_void_p(stream_ptr or 0)))

    def set_frame_times(self, cap: Sequence[float], exp: Sequence[float]):
        cap = np.ascontiguousarray(cap, dtype=np.float64)
        exp = np.ascontiguousarray(exp, dtype=np.float64)
        self._check(self.lib.mbavo_set_frame_times(self._h, C.c_int(cap.shape[0]), _dp(cap), _dp(exp)))

    def set_level(self, level: int, lv, point_slice: Optional[slice] = None):
        """Upload one synth.Level (host memory).  point_slice selects a shard of the host-map points."""
        xy = lv.xy if point_slice is None else np.ascontiguousarray(lv.xy[point_slice])
        z = lv.z if point_slice is None else np.ascontiguousarray(lv.z[point_slice])
        F = len(lv.cur_I)
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in lv.cur_I])
        d = _Level(MEM_HOST, lv.H, lv.W, lv.fx, lv.fy, lv.cx, lv.cy, lv.ref_I.ctypes.data, lv.ref_dIxy.ctypes.data, cur, F,
                   xy.ctypes.data, 16, 0, z.ctypes.data, xy.shape[0], lv.pattern.ctypes.data, lv.S, lv.N, None, None, 0)
        self._check(self.lib.mbavo_set_level(self._h, C.c_int(level), C.byref(d)))

    def set_level_device(self, level: int, H, W, fx, fy, cx, cy, ref_I_ptr, ref_dIxy_ptr, cur_I_ptrs, xy_ptr, xy_stride,
                         xy_offset, z_ptr, P, pattern_ptr, S, N, ext_flags_ptr=None, ext_patch_cost_ptr=None,
                         ext_patch_cost_stride=0):
        """Device-pointer variant (pointers stay owned by the caller, e.g. torch tensors)."""
        F = len(cur_I_ptrs)
        cur = (C.c_void_p * F)(*cur_I_ptrs)
        d = _Level(MEM_DEVICE, H, W, fx, fy, cx, cy, ref_I_ptr, ref_dIxy_ptr, cur, F, xy_ptr, xy_stride, xy_offset, z_ptr, P,
                   pattern_ptr, S, N, ext_flags_ptr, ext_patch_cost_ptr, ext_patch_cost_stride)
        self._check(self.lib.mbavo_set_level(self._h, C.c_int(level), C.byref(d)))

    def set_keyframe_pyramid(self, n_levels: int, ref_I0: np.ndarray):
        """mbavo_set_keyframe_pyramid: level-0 keyframe (host uint8); coarser levels, gradients, texels built on the GPU."""
        H0, W0 = ref_I0.shape
        self._check(self.lib.mbavo_set_keyframe_pyramid(self._h, C.c_int(n_levels), C.c_int(MEM_HOST), C.c_void_p(ref_I0.ctypes.data),
                                                        C.c_int(H0), C.c_int(W0)))

    def set_live_pyramid(self, n_levels: int, cur_I0: Sequence[np.ndarray]):
        F = len(cur_I0)
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in cur_I0])
        self._check(self.lib.mbavo_set_live_pyramid(self._h, C.c_int(n_levels), C.c_int(MEM_HOST), cur, C.c_int(F)))

    def set_level_points(self, level: int, lv, point_slice: Optional[slice] = None):
        """mbavo_set_level_points from a synth.Level: intrinsics, host-map points, pattern, exposure samples."""
        xy = lv.xy if point_slice is None else np.ascontiguousarray(lv.xy[point_slice])
        z = lv.z if point_slice is None else np.ascontiguousarray(lv.z[point_slice])
        d = _LevelPoints(MEM_HOST, lv.fx, lv.fy, lv.cx, lv.cy, xy.ctypes.data, 16, 0, z.ctypes.data, xy.shape[0],
                         lv.pattern.ctypes.data, lv.S, lv.N)
        self._check(self.lib.mbavo_set_level_points(self._h, C.c_int(level), C.byref(d)))

    def set_points_pyramid(self, levels: Sequence):
        """mbavo_set_points_pyramid from a list of synth.Level (level 0 first)."""
        arr = (_LevelPoints * len(levels))()
        for l, lv in enumerate(levels):
            arr[l] = _LevelPoints(MEM_HOST, lv.fx, lv.fy, lv.cx, lv.cy, lv.xy.ctypes.data, 16, 0, lv.z.ctypes.data, lv.xy.shape[0],
                                  lv.pattern.ctypes.data, lv.S, lv.N)
        self._check(self.lib.mbavo_set_points_pyramid(self._h, C.c_int(len(levels)), arr))

    def set_frame(self, n_levels: int, ref_I0: Optional[np.ndarray], cur_I0: Optional[Sequence[np.ndarray]], levels: Sequence,
                  async_upload: bool = False):
        """mbavo_set_frame: level-0 keyframe + live images (either may be None) and the points of every level (synth.Level list,
        level 0 first) in one call.  async_upload: no synchronisation — the arrays must stay alive and unchanged until the next
        blocking call on this context returns (they are kept referenced here)."""
        arr = (_LevelPoints * len(levels))()
        for l, lv in enumerate(levels):
            arr[l] = _LevelPoints(MEM_HOST, lv.fx, lv.fy, lv.cx, lv.cy, lv.xy.ctypes.data, 16, 0, lv.z.ctypes.data, lv.xy.shape[0],
                                  lv.pattern.ctypes.data, lv.S, lv.N)
        H0, W0 = ref_I0.shape if ref_I0 is not None else (0, 0)
        F = len(cur_I0) if cur_I0 is not None else 0
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in cur_I0]) if F else None
        self._keep["frame"] = (ref_I0, cur_I0, levels, arr, cur)
        self._check(self.lib.mbavo_set_frame(self._h, C.c_int(n_levels), C.c_int(MEM_HOST),
                                             C.c_void_p(ref_I0.ctypes.data if ref_I0 is not None else 0), C.c_int(H0), C.c_int(W0), cur,
                                             C.c_int(F), arr, C.c_int(1 if async_upload else 0)))

    def prepare_frame(self, n_levels: int, ref_I0: Optional[np.ndarray], cur_I0: Optional[Sequence[np.ndarray]], levels: Sequence,
                      async_upload: bool = False):
        """set_frame with its argument marshalling done ONCE: returns a callable that only makes the C call (a caller that re-uses
        its frame buffers — pinned staging memory — pays the ctypes set-up of ~20 structure fields per level once, not per frame;
        a C / C++ caller of mbavo_set_frame never pays it).  The arrays are kept referenced by the returned callable."""
        arr = (_LevelPoints * len(levels))()
        for l, lv in enumerate(levels):
            arr[l] = _LevelPoints(MEM_HOST, lv.fx, lv.fy, lv.cx, lv.cy, lv.xy.ctypes.data, 16, 0, lv.z.ctypes.data, lv.xy.shape[0],
                                  lv.pattern.ctypes.data, lv.S, lv.N)
        H0, W0 = ref_I0.shape if ref_I0 is not None else (0, 0)
        F = len(cur_I0) if cur_I0 is not None else 0
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in cur_I0]) if F else None
        keep = (ref_I0, cur_I0, levels, arr, cur)
        args = (self._h, C.c_int(n_levels), C.c_int(MEM_HOST), C.c_void_p(ref_I0.ctypes.data if ref_I0 is not None else 0), C.c_int(H0),
                C.c_int(W0), cur, C.c_int(F), arr, C.c_int(1 if async_upload else 0))
        fn, check = self.lib.mbavo_set_frame, self._check

        def call(_keep=keep):
            check(fn(*args))

        return call

    def prepare_gn_sweep(self, level_coarse: int, level_fine: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float,
                         radius: float = 1e4, chain: bool = False, solver_type: int = SOLVER_SVD_JACOBI):
        """gn_sweep with its argument marshalling done once: returns a callable that restores the starting knots into its own
        buffers, makes the C call and returns (costs, knots_t, knots_R) — views of buffers that the next call overwrites."""
        kt0 = np.array(knots_t, dtype=np.float64, order="C")
        kR0 = np.array(knots_R, dtype=np.float64, order="C")
        kt, kR = kt0.copy(), kR0.copy()
        costs = np.zeros((level_coarse - level_fine + 1, 2))
        args = (self._h, C.c_int(level_coarse), C.c_int(level_fine), C.c_int(1 if chain else 0), C.c_int(k), C.c_double(t0), C.c_double(dt),
                C.c_int(kt.shape[0]), _dp(kt), _dp(kR), C.c_double(radius), C.c_double(huber_a), C.c_int(solver_type), _dp(costs))
        fn, check, copyto = self.lib.mbavo_gn_sweep, self._check, np.copyto

        def call():
            copyto(kt, kt0)
            copyto(kR, kR0)
            check(fn(*args))
            return costs, kt, kR

        return call

    def set_live_images(self, level: int, cur_I: Sequence[np.ndarray]):
        """mbavo_set_live_images with host images: a new blurred frame for a level whose keyframe stays resident."""
        F = len(cur_I)
        cur = (C.c_void_p * F)(*[c.ctypes.data for c in cur_I])
        self._check(self.lib.mbavo_set_live_images(self._h, C.c_int(level), C.c_int(MEM_HOST), cur, C.c_int(F)))

    def set_outliers(self, level: int, flags: Optional[np.ndarray], num_bad: int = 0):
        if flags is None:
            self._check(self.lib.mbavo_set_outliers(self._h, C.c_int(level), None, C.c_int(0)))
        else:
            fl = np.ascontiguousarray(flags, dtype=np.uint8)
            self._check(self.lib.mbavo_set_outliers(self._h, C.c_int(level), fl.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                                    C.c_int(num_bad)))

    def set_num_bad(self, level: int, num_bad: int):
        self._check(self.lib.mbavo_set_num_bad(self._h, C.c_int(level), C.c_int(num_bad)))

    # -- evaluation ---------------------------------------------------------------------------------------
    @staticmethod
    def _spline(k, t0, dt, knots_t, knots_R):
        kt = np.ascontiguousarray(knots_t, dtype=np.float64).reshape(-1)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64).reshape(-1)
        n = kt.shape[0] // 3
        return _Spline(k, t0, dt, n, _dp(kt), _dp(kR)), kt, kR, n

    def evaluate(self, level: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float,
                 with_hessian: bool = True):
        """mbavo_evaluate -> (cost, H, g); H, g are None for the cost-only branch."""
        sp, kt, kR, n = self._spline(k, t0, dt, knots_t, knots_R)
        cost = C.c_double(0)
        if with_hessian:
            H = np.zeros((6 * n, 6 * n))
            g = np.zeros(6 * n)
            self._check(self.lib.mbavo_evaluate(self._h, C.c_int(level), C.byref(sp), C.c_double(huber_a), C.byref(cost),
                                                _dp(H), _dp(g)))
            return cost.value, H, g
        self._check(self.lib.mbavo_evaluate(self._h, C.c_int(level), C.byref(sp), C.c_double(huber_a), C.byref(cost), None,
                                            None))
        return cost.value, None, None

    def evaluate_async(self, level: int, k, t0, dt, knots_t, knots_R, huber_a, with_hessian: bool,
                       num_residuals_global: int, packed_dev_ptr: int):
        """mbavo_evaluate_async -> (kmin, knot_window); the packed vector is left at packed_dev_ptr (device)."""
        sp, kt, kR, n = self._spline(k, t0, dt, knots_t, knots_R)
        kmin, nk = C.c_int(0), C.c_int(0)
        self._check(self.lib.mbavo_evaluate_async(self._h, C.c_int(level), C.byref(sp), C.c_double(huber_a),
                                                  C.c_int(1 if with_hessian else 0), C.c_longlong(num_residuals_global),
                                                  C.c_void_p(packed_dev_ptr), C.byref(kmin), C.byref(nk)))
        return kmin.value, nk.value

    def debug_dump(self, level: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float, F: int, N: int, P: int, S: int):
        """mbavo_debug_dump -> dict of the per-stage intermediates of one Hessian-pass evaluation (see include/mbavo.h)."""
        sp, kt, kR, n = self._spline(k, t0, dt, knots_t, knots_R)
        poses, wts, theta = np.zeros((F, N, 7)), np.zeros((F, N, k)), np.zeros((F, N, k, 3, 3))
        seg = np.zeros((F, N), dtype=np.int32)
        centres = np.zeros((F, P, 2))
        r = np.zeros((F, P, S), dtype=np.float32)
        J = np.zeros(F * P * S * 6 * 8, dtype=np.float32)  # worst-case window
        out = _DebugOut(poses.ctypes.data, wts.ctypes.data, theta.ctypes.data, seg.ctypes.data, centres.ctypes.data, r.ctypes.data,
                        J.ctypes.data, 0, 0, 0, 0.0)
        self._check(self.lib.mbavo_debug_dump(self._h, C.c_int(level), C.byref(sp), C.c_double(huber_a), C.byref(out)))
        d = 6 * out.knot_window
        return dict(poses_tq=poses, blend_weights=wts, theta=theta, segment_start_knot=seg, patch_centres=centres, residuals=r,
                    jacobians=J[: F * P * S * d].reshape(F, P, S, d).copy(), kmin=out.kmin, knot_window=out.knot_window, cost=out.cost)

    def cost_tma(self, level: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float, box_w: int, box_h: int):
        """mbavo_debug_cost_tma -> (cost, kernel_ms, fraction of samples served from the TMA tiles)."""
        sp, kt, kR, n = self._spline(k, t0, dt, knots_t, knots_R)
        cost, ms, frac = C.c_double(0), C.c_float(0), C.c_double(0)
        self._check(self.lib.mbavo_debug_cost_tma(self._h, C.c_int(level), C.byref(sp), C.c_double(huber_a), C.c_int(box_w), C.c_int(box_h),
                                                  C.byref(cost), C.byref(ms), C.byref(frac)))
        return cost.value, ms.value, frac.value

    def packed_len(self, knot_window: int) -> int:
        return int(self.lib.mbavo_packed_len(C.c_int(knot_window)))

    def unpack(self, packed: np.ndarray, kmin: int, knot_window: int, n_knots: int, with_hessian: bool = True):
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        cost = C.c_double(0)
        H = np.zeros((6 * n_knots, 6 * n_knots)) if with_hessian else None
        g = np.zeros(6 * n_knots) if with_hessian else None
        self._check(self.lib.mbavo_unpack(_dp(packed), C.c_int(kmin), C.c_int(knot_window), C.c_int(n_knots), C.byref(cost),
                                          _dp(H) if with_hessian else None, _dp(g) if with_hessian else None))
        return cost.value, H, g

    def patch_costs(self, level: int, n_frames: int, num_keypoints: int) -> np.ndarray:
        out = np.zeros((n_frames, num_keypoints))
        self._check(self.lib.mbavo_patch_costs(self._h, C.c_int(level), _dp(out)))
        return out

    def detect_outliers(self, level: int, max_chi_square_error: float) -> int:
        nb = C.c_int(0)
        self._check(self.lib.mbavo_detect_outliers(self._h, C.c_int(level), C.c_double(max_chi_square_error), C.byref(nb)))
        return nb.value

    # -- host-side solver -----------------------------------------------------------------------------------
    def trust_region_step(self, H: np.ndarray, g: np.ndarray, radius: float, solver_type: int = SOLVER_SVD_JACOBI):
        """Damps H in place (as the reference does) and returns (step, model_cost_change)."""
        assert H.flags.c_contiguous and H.dtype == np.float64
        g = np.ascontiguousarray(g, dtype=np.float64)
        step = np.zeros_like(g)
        model = C.c_double(0)
        self._check(self.lib.mbavo_trust_region_step(_dp(H), _dp(g), C.c_int(g.shape[0]), C.c_double(radius),
                                                     C.c_int(solver_type), _dp(step), C.byref(model)))
        return step, model.value

    def spline_plus(self, knots_t, knots_R, step):
        kt = np.ascontiguousarray(knots_t, dtype=np.float64)
        kR = np.ascontiguousarray(knots_R, dtype=np.float64)
        st = np.ascontiguousarray(step, dtype=np.float64)
        n = kt.shape[0]
        ct, cR = np.zeros_like(kt), np.zeros_like(kR)
        self._check(self.lib.mbavo_spline_plus(C.c_int(n), _dp(kt), _dp(kR), _dp(st), _dp(ct), _dp(cR)))
        return ct, cR

    def gn_iteration(self, level: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float, radius: float = 1e4,
                     solver_type: int = SOLVER_SVD_JACOBI):
        """mbavo_gn_iteration -> (cost, candidate_cost, step, cand_t, cand_R)."""
        sp, kt, kR, n = self._spline(k, t0, dt, knots_t, knots_R)
        cost, cand = C.c_double(0), C.c_double(0)
        step, ct, cR = np.zeros(6 * n), np.zeros((n, 3)), np.zeros((n, 4))
        self._check(self.lib.mbavo_gn_iteration(self._h, C.c_int(level), C.c_int(k), C.c_double(t0), C.c_double(dt), C.c_int(n),
                                                _dp(kt), _dp(kR), C.c_double(radius), C.c_double(huber_a), C.c_int(solver_type),
                                                C.byref(cost), C.byref(cand), _dp(step), _dp(ct), _dp(cR)))
        return cost.value, cand.value, step, ct, cR

    def gn_sweep(self, level_coarse: int, level_fine: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float,
                 radius: float = 1e4, chain: bool = False, solver_type: int = SOLVER_SVD_JACOBI):
        """mbavo_gn_sweep -> (costs [(cost, candidate cost) per level, coarse first], knots_t, knots_R)."""
        kt = np.array(knots_t, dtype=np.float64, order="C")
        kR = np.array(knots_R, dtype=np.float64, order="C")
        n = kt.shape[0]
        costs = np.zeros((level_coarse - level_fine + 1, 2))
        self._check(self.lib.mbavo_gn_sweep(self._h, C.c_int(level_coarse), C.c_int(level_fine), C.c_int(1 if chain else 0), C.c_int(k),
                                            C.c_double(t0), C.c_double(dt), C.c_int(n), _dp(kt), _dp(kR), C.c_double(radius),
                                            C.c_double(huber_a), C.c_int(solver_type), _dp(costs)))
        return costs, kt, kR

    def optimize_level(self, level: int, k: int, t0: float, dt: float, knots_t, knots_R, huber_a: float = 10.0,
                       max_chi_square_error: float = 3.0, solver_type: int = SOLVER_SVD_JACOBI, **overrides):
        """mbavo_optimize_level -> (knots_t, knots_R, summary dict)."""
        kt = np.array(knots_t, dtype=np.float64, order="C")
        kR = np.array(knots_R, dtype=np.float64, order="C")
        n = kt.shape[0]
        opt = _LmOptions()
        self.lib.mbavo_lm_default_options(C.byref(opt))
        opt.huber_a, opt.max_chi_square_error, opt.solver_type = huber_a, max_chi_square_error, solver_type
        for k_, v in overrides.items():
            setattr(opt, k_, v)
        summ = _LmSummary()
        self._check(self.lib.mbavo_optimize_level(self._h, C.c_int(level), C.c_int(k), C.c_double(t0), C.c_double(dt),
                                                  C.c_int(n), _dp(kt), _dp(kR), C.byref(opt), C.byref(summ)))
        s = {f: getattr(summ, f) for f, _ in _LmSummary._fields_ if f not in ("first_step", "decisions")}
        s["first_step"] = np.array(summ.first_step[: 6 * n])
        s["decisions"] = summ.decisions.decode()
        return kt, kR, s

    # -- point sharding ----------------------------------------------------------------------------------------
    def shard_export(self):
        """mbavo_shard_export -> (ipc_handle bytes, mailbox device pointer)."""
        h = (C.c_ubyte * IPC_HANDLE_BYTES)()
        ptr = C.c_void_p()
        self._check(self.lib.mbavo_shard_export(self._h, h, C.byref(ptr)))
        return bytes(h), int(ptr.value)

    def shard_connect(self, world: int, rank: int, handles: Optional[Sequence[bytes]] = None,
                      mailbox_ptrs: Optional[Sequence[int]] = None):
        """mbavo_shard_connect: `handles` (other processes) or `mailbox_ptrs` (same process), both in rank order."""
        hb = None
        if handles is not None:
            hb = (C.c_ubyte * (IPC_HANDLE_BYTES * world)).from_buffer_copy(b"".join(handles))
        pp = (C.c_void_p * world)(*mailbox_ptrs) if mailbox_ptrs is not None else None
        self._check(self.lib.mbavo_shard_connect(self._h, C.c_int(world), C.c_int(rank), hb, pp))

    def shard_disconnect(self):
        self._check(self.lib.mbavo_shard_disconnect(self._h))

    def shard_set_global_points(self, level: int, num_keypoints_global: int):
        self._check(self.lib.mbavo_shard_set_global_points(self._h, C.c_int(level), C.c_int(num_keypoints_global)))

    def keyframe_stats(self, level: int, poses_tq) -> tuple:
        """mbavo_keyframe_stats -> (avg_flow, avg_kernel_len); poses_tq is (3, 7): T0, T-, T+ as tx ty tz qx qy qz qw."""
        poses = np.ascontiguousarray(poses_tq, dtype=np.float64).reshape(3, 7)
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.mbavo_keyframe_stats(self._h, C.c_int(level), _dp(poses), C.byref(a), C.byref(b)))
        return a.value, b.value

    def select_points(self, n_levels: int, depth_z: np.ndarray, fx: float, fy: float, cx: float, cy: float, pattern: np.ndarray,
                      num_virtual_poses: int, score_threshold: float = 25.0, cell_H: int = 30, cell_W: int = 30):
        """mbavo_select_points (semi-dense selection + grid selection + depth look-up on the GPU) -> points per level."""
        depth = np.ascontiguousarray(depth_z, dtype=np.float32)
        pat = np.ascontiguousarray(pattern, dtype=np.int32)
        sel = _PointSelection(score_threshold, cell_H, cell_W, MEM_HOST, depth.ctypes.data, fx, fy, cx, cy, pat.ctypes.data,
                              pat.shape[0], num_virtual_poses)
        counts = (C.c_int * n_levels)()
        self._check(self.lib.mbavo_select_points(self._h, C.c_int(n_levels), C.byref(sel), counts))
        return [int(c) for c in counts]

    def get_points(self, level: int):
        """mbavo_get_points -> (xy (P, 2), z (P,)) of a level, copied from the device."""
        n = C.c_int(0)
        self._check(self.lib.mbavo_get_points(self._h, C.c_int(level), C.c_int(0), None, None, C.byref(n)))
        xy = np.zeros((n.value, 2))
        z = np.zeros(n.value)
        if n.value:
            self._check(self.lib.mbavo_get_points(self._h, C.c_int(level), C.c_int(n.value), _dp(xy), _dp(z), C.byref(n)))
        return xy, z

    # -- introspection ----------------------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        return int(self.lib.mbavo_kernel_launches(self._h))

    def device_sweeps(self) -> int:
        return int(self.lib.mbavo_device_sweeps(self._h))

    def persistent_sweeps(self) -> int:
        return int(self.lib.mbavo_persistent_sweeps(self._h))

    def sweep_pass_times(self) -> np.ndarray:
        """mbavo_sweep_pass_times -> microseconds per pass of the last persistent sweep, shape (levels, 2): [Hessian pass, cost pass]."""
        out = np.zeros(2 * MAX_LEVELS)
        n = C.c_int(0)
        self._check(self.lib.mbavo_sweep_pass_times(self._h, _dp(out), C.c_int(out.shape[0]), C.byref(n)))
        return out[: n.value].reshape(-1, 2).copy()

    def level_uses_texels(self, level: int) -> int:
        return int(self.lib.mbavo_level_uses_texels(self._h, C.c_int(level)))

    def enable_kernel_timing(self, on: bool):
        self._check(self.lib.mbavo_enable_kernel_timing(self._h, C.c_int(1 if on else 0)))

    def last_kernel_ms(self) -> float:
        return float(self.lib.mbavo_last_kernel_ms(self._h))


def synthesize_blurred(ref_I: np.ndarray, plane_depth: float, fx: float, fy: float, cx: float, cy: float,
                       poses_tq: np.ndarray, device: int = -1) -> np.ndarray:
    """mbavo_synthesize_blurred with host images: poses_tq is (N, 7) = tx ty tz qx qy qz qw per exposure sample."""
    lib = load_library()
    ref_I = np.ascontiguousarray(ref_I, dtype=np.uint8)
    poses = np.ascontiguousarray(poses_tq, dtype=np.float64).reshape(-1, 7)
    H, W = ref_I.shape
    out = np.zeros((H, W), dtype=np.uint8)
    rc = lib.mbavo_synthesize_blurred(C.c_int(device), C.c_int(MEM_HOST), C.c_void_p(ref_I.ctypes.data), C.c_int(H), C.c_int(W),
                                      C.c_double(plane_depth), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                      _dp(poses), C.c_int(poses.shape[0]), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise MbavoError(f"mbavo_synthesize_blurred failed: {rc}")
    return out


def limits_for(prob, n_frames: Optional[int] = None) -> Limits:
    return Limits(max_num_frames=n_frames or prob.F,
                  max_num_virtual_poses_per_frame=max(lv.N for lv in prob.levels),
                  max_num_keypoints=max(lv.P for lv in prob.levels),
                  max_patch_size=max(lv.S for lv in prob.levels),
                  max_num_ctrl_knots=max(prob.n_knots, 2))


def upload_problem(ctx: Context, prob) -> None:
    ctx.set_frame_times(prob.cap, prob.exp)
    for l, lv in enumerate(prob.levels):
        ctx.set_level(l, lv)


class FrameTracker:
    """mbavo_tracker + mbavo_track_frame: BlurAwareDirectTracker::trackFrame on a context that holds the keyframe and points."""

    def __init__(self, ctx: "Context", n_levels: int, sample_dt: float, keyframe_capture_time: float, spline_deg_k: int = 2, **lm):
        self.ctx, self.n_levels = ctx, n_levels
        self.state = _Tracker()
        _chk(ctx.lib.mbavo_tracker_init(C.byref(self.state), C.c_int(spline_deg_k), C.c_double(sample_dt), C.c_double(keyframe_capture_time)))
        self.opt = _LmOptions()
        ctx.lib.mbavo_lm_default_options(C.byref(self.opt))
        for k, v in lm.items():
            setattr(self.opt, k, v)

    def track(self, cur_I0: np.ndarray, capture_time: float, exposure_time: float) -> dict:
        img = np.ascontiguousarray(cur_I0, dtype=np.uint8)
        res = _FrameResult()
        self.ctx._check(self.ctx.lib.mbavo_track_frame(self.ctx._h, C.byref(self.state), C.c_int(self.n_levels), C.c_int(MEM_HOST),
                                                       C.c_void_p(img.ctypes.data), C.c_double(capture_time), C.c_double(exposure_time),
                                                       C.byref(self.opt), C.byref(res)))
        return dict(t_cur2key=np.array(res.t_cur2key), q_cur2key=np.array(res.q_cur2key), t_cur2world=np.array(res.t_cur2world),
                    q_cur2world=np.array(res.q_cur2world), avg_flow=res.avg_flow, avg_kernel_len=res.avg_kernel_len,
                    levels_run=res.levels_run,
                    levels=[{f: getattr(res.levels[l], f) for f, _ in _LmSummary._fields_ if f not in ("first_step",)}
                            for l in range(self.n_levels)])

    def new_keyframe(self, capture_time: float):
        _chk(self.ctx.lib.mbavo_tracker_new_keyframe(C.byref(self.state), C.c_double(capture_time)))

    @property
    def knots(self):
        n = self.state.num_ctrl_knots
        return np.array(self.state.knots_t[:3 * n]).reshape(n, 3), np.array(self.state.knots_R[:4 * n]).reshape(n, 4)

    @property
    def velocity(self):
        return np.array(self.state.velocity)


# -- per-frame trajectory bookkeeping (host functions of the library; no context, no GPU) --------------------------------
def _chk(rc: int):
    if rc != 0:
        raise MbavoError(f"mbavo error {rc}")


def se3_exp(tangent):
    tg = np.ascontiguousarray(tangent, dtype=np.float64)
    t, q = np.zeros(3), np.zeros(4)
    _chk(load_library().mbavo_se3_exp(_dp(tg), _dp(t), _dp(q)))
    return t, q


def se3_log(t, q):
    t, q = np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros(6)
    _chk(load_library().mbavo_se3_log(_dp(t), _dp(q), _dp(out)))
    return out


def spline_pose(k, t0, dt, knots_t, knots_R, time):
    sp, kt, kR, n = Context._spline(k, t0, dt, knots_t, knots_R)
    t, q = np.zeros(3), np.zeros(4)
    _chk(load_library().mbavo_spline_pose(C.byref(sp), C.c_double(time), _dp(t), _dp(q)))
    return t, q


def spline_transform_to(k, t0, dt, knots_t, knots_R, time, target_t, target_q):
    sp, kt, kR, n = Context._spline(k, t0, dt, knots_t, knots_R)
    tt, tq = np.ascontiguousarray(target_t, dtype=np.float64), np.ascontiguousarray(target_q, dtype=np.float64)
    ot, oR = np.zeros((n, 3)), np.zeros((n, 4))
    _chk(load_library().mbavo_spline_transform_to(C.byref(sp), C.c_double(time), _dp(tt), _dp(tq), _dp(ot), _dp(oR)))
    return ot, oR


def predict_spline(knots_t, knots_R, velocity, dt_frame):
    kt = np.array(knots_t, dtype=np.float64, order="C")
    kR = np.array(knots_R, dtype=np.float64, order="C")
    v = np.ascontiguousarray(velocity, dtype=np.float64)
    _chk(load_library().mbavo_predict_spline(C.c_int(kt.shape[0]), _dp(kt), _dp(kR), _dp(v), C.c_double(dt_frame)))
    return kt, kR


def frame_velocity(prev_t, prev_q, cur_t, cur_q, dt_frame):
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (prev_t, prev_q, cur_t, cur_q)]
    out = np.zeros(6)
    _chk(load_library().mbavo_frame_velocity(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), C.c_double(dt_frame), _dp(out)))
    return out


def upload_problem_pyramid(ctx: Context, prob) -> None:
    """The same problem through the device-built pyramid path: only the level-0 images travel."""
    ctx.set_frame_times(prob.cap, prob.exp)
    n = len(prob.levels)
    ctx.set_keyframe_pyramid(n, prob.levels[0].ref_I)
    ctx.set_live_pyramid(n, prob.levels[0].cur_I)
    ctx.set_points_pyramid(prob.levels)


def optimize_trajectory(ctx: Context, prob, **kw):
    """BlurAwareDirectTracker::optimizeTrajectory (blur_aware_direct_tracker.cpp:544-588): coarse -> fine over the
    pyramid levels already uploaded into ctx."""
    kt, kR = prob.knots_t.copy(), prob.knots_R.copy()
    summaries: List[dict] = []
    for level in reversed(range(len(prob.levels))):
        kt, kR, s = ctx.optimize_level(level, prob.k, prob.t0, prob.dt, kt, kR, huber_a=prob.huber_a,
                                       max_chi_square_error=prob.max_chi_square_error, **kw)
        summaries.append(s)
    return kt, kR, summaries
