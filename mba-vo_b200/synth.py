"""Seeded synthetic inputs for the blur-aware tracking hot path.

Everything the path consumes is generated here with numpy only: keyframe texture, image pyramid
(2x2 box, truncating cast — src/core/measurements/ImagePyramid.h:59-99), gradient images (0.5 * central
difference, zero border, interleaved — src/core/image_proc/Gradient.h:17-75), host-map points with depth,
the 8-pixel residual pattern of the reference test (test/test_blur_aware_tracker_modules.cpp:662-679), a
ground-truth SE(3) spline, the live (blurred) frame synthesised as in
src/ba_tracker/generate_synthetic_data.cpp:127-180, and perturbed initial knots.

The five named configs are BASELINE.json's (`C1`..`C5`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# residual pattern of the reference test, (dx, dy) pairs
PATTERN8 = np.array([[-2, -2], [2, -2], [-1, -1], [1, -1], [0, 0], [0, 1], [-2, 2], [2, 2]], dtype=np.int32)


@dataclass
class Level:
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float
    ref_I: np.ndarray        # (H, W) uint8      keyframe image
    ref_dIxy: np.ndarray     # (H, W, 2) float32 keyframe gradient, interleaved (dx, dy)
    cur_I: List[np.ndarray]  # F x (H, W) uint8  live (blurred) frames
    xy: np.ndarray           # (P, 2) float64    host-map points (keyframe pixel coordinates)
    z: np.ndarray            # (P,) float64      depth of every point
    pattern: np.ndarray      # (S, 2) int32
    N: int                   # exposure samples (virtual poses) per frame

    @property
    def P(self) -> int:
        return int(self.xy.shape[0])

    @property
    def S(self) -> int:
        return int(self.pattern.shape[0])


@dataclass
class Problem:
    name: str
    levels: List[Level]
    cap: np.ndarray          # (F,) capture times
    exp: np.ndarray          # (F,) exposure times
    k: int                   # spline order (control knots per segment): 2 linear, 4 cubic
    t0: float                # spline start time
    dt: float                # knot spacing
    knots_t: np.ndarray      # (n, 3) initial control-knot translations
    knots_R: np.ndarray      # (n, 4) initial control-knot rotations, quaternion (x, y, z, w)
    gt_knots_t: np.ndarray
    gt_knots_R: np.ndarray
    huber_a: float = 10.0
    max_chi_square_error: float = 3.0
    seg_start: np.ndarray = field(default_factory=lambda: np.zeros(1, dtype=np.int32))

    @property
    def F(self) -> int:
        return int(self.cap.shape[0])

    @property
    def n_knots(self) -> int:
        return int(self.knots_t.shape[0])


# ---------------------------------------------------------------------------------------------------------
# quaternion / spline helpers (x, y, z, w); only used to build inputs
# ---------------------------------------------------------------------------------------------------------
def q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def q_conj(a):
    return np.array([-a[0], -a[1], -a[2], a[3]])


def q_exp(phi):
    phi = np.asarray(phi, dtype=np.float64)
    th = np.linalg.norm(phi)
    if th * th < 1e-20:
        return np.array([0.5 * phi[0], 0.5 * phi[1], 0.5 * phi[2], 1.0])
    return np.concatenate([np.sin(0.5 * th) / th * phi, [np.cos(0.5 * th)]])


def q_log(q):
    v = np.asarray(q[:3], dtype=np.float64)
    n = np.linalg.norm(v)
    if n * n < 1e-20:
        return 2.0 / q[3] * v
    return 2.0 * np.arctan(n / q[3]) / n * v


def q_to_R(q):
    x, y, z, w = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def spline_pose(k, knots_t, knots_R, t0, dt, t):
    """Pose (t, q) on the spline at time t; k=2 linear, k=4 cumulative cubic B-spline."""
    s = (t - t0) / dt
    idx = min(int(s), len(knots_t) - k)  # the end of the last segment belongs to it (u = 1)
    u = s - idx
    kt, kR = knots_t[idx:idx + k], knots_R[idx:idx + k]
    if k == 2:
        wt, wr = [1 - u, u], [1, u]
    else:
        uu, uuu, s6 = u * u, u * u * u, 1.0 / 6.0
        wt = [s6 - 0.5 * u + 0.5 * uu - s6 * uuu, 4 * s6 - uu + 0.5 * uuu, s6 + 0.5 * u + 0.5 * uu - 0.5 * uuu, s6 * uuu]
        wr = [1, 5 * s6 + 0.5 * u - 0.5 * uu + s6 * uuu, s6 + 0.5 * u + 0.5 * uu - 2 * s6 * uuu, s6 * uuu]
    tt = sum(w * p for w, p in zip(wt, kt))
    q = kR[0]
    for j in range(1, k):
        q = q_mul(q, q_exp(wr[j] * q_log(q_mul(q_conj(kR[j - 1]), kR[j]))))
    return tt, q


# ---------------------------------------------------------------------------------------------------------
# images
# ---------------------------------------------------------------------------------------------------------
def make_texture(H: int, W: int, seed: int) -> np.ndarray:
    """8-bit band-limited texture: a sum of seeded random sinusoids, gradients non-degenerate everywhere."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    img = np.zeros((H, W))
    for _ in range(24):
        wavelength = rng.uniform(12.0, 90.0)
        ang = rng.uniform(0, 2 * np.pi)
        ph = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(0.4, 1.0) * (wavelength / 90.0) ** 0.5
        img += amp * np.sin(2 * np.pi / wavelength * (np.cos(ang) * xx + np.sin(ang) * yy) + ph)
    img = (img - img.min()) / (img.max() - img.min())
    return np.clip(np.rint(20 + 215 * img), 0, 255).astype(np.uint8)


def ramp_image(H: int, W: int) -> np.ndarray:
    """The reference test's image, (c + r) % 255 (test/test_blur_aware_tracker_modules.cpp:69-81)."""
    r, c = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return ((c + r) % 255).astype(np.uint8)


# the reference's synthetic scene: five rectangles (top-left corner, size) and two triangles, white on black
SHAPES_RECTS = [((300, 50), (50, 100)), ((250, 200), (100, 50)), ((400, 300), (100, 100)), ((500, 50), (100, 100)), ((250, 300), (100, 100))]
SHAPES_TRIANGLES = [[(500, 50), (400, 150), (550, 250)], [(150, 300), (50, 450), (250, 400)]]


def _line_pixels(p0, p1):
    """cv::LineIterator, 8-connected, left to right (what cv::fillPoly draws along every polygon edge with LINE_8)."""
    if p1[0] < p0[0]:
        p0, p1 = p1, p0
    (x, y), (x1, y1) = p0, p1
    dx, dy = x1 - x, abs(y1 - y)
    sy = 1 if y1 >= y else -1
    out = []
    if dx >= dy:
        err = dx - 2 * dy
        for _ in range(dx + 1):
            out.append((x, y))
            if err < 0:
                err += 2 * dx
                y += sy
            err -= 2 * dy
            x += 1
    else:
        err = dy - 2 * dx
        for _ in range(dy + 1):
            out.append((x, y))
            if err < 0:
                err += 2 * dy
                x += 1
            err -= 2 * dx
            y += sy
    return out


def shapes_image(H: int = 480, W: int = 640) -> np.ndarray:
    """synthesize_img_with_rand_shapes (src/ba_tracker/generate_synthetic_data.cpp:11-125): the reference's own test scene — five
    rectangles and two triangles filled white (255) on black by cv::fillPoly(..., LINE_8).  For these integer-vertex convex
    polygons fillPoly covers the lattice points of the closed polygon plus the Bresenham line (cv::LineIterator, 8-connected,
    left to right) along every edge; pinned byte for byte to OpenCV's own output in tests/golden/shapes.npz."""
    im = np.zeros((H, W), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    polys = [[(x, y), (x + w, y), (x + w, y + h), (x, y + h)] for (x, y), (w, h) in SHAPES_RECTS] + SHAPES_TRIANGLES
    for pts in polys:
        n = len(pts)
        area = sum(pts[i][0] * pts[(i + 1) % n][1] - pts[(i + 1) % n][0] * pts[i][1] for i in range(n))
        sgn = 1 if area > 0 else -1
        inside = np.ones((H, W), dtype=bool)
        for i in range(n):
            (x0, y0), (x1, y1) = pts[i], pts[(i + 1) % n]
            inside &= sgn * ((x1 - x0) * (yy - y0) - (y1 - y0) * (xx - x0)) >= 0
        im[inside] = 255
        for i in range(n):
            for x, y in _line_pixels(pts[i], pts[(i + 1) % n]):
                if 0 <= x < W and 0 <= y < H:
                    im[y, x] = 255
    return im


def image_gradient(I: np.ndarray) -> np.ndarray:
    """Gradient.h:17-75 — 0.5 * central difference, zero 1-px border, interleaved (dx, dy) float32."""
    H, W = I.shape
    f = I.astype(np.float32)
    g = np.zeros((H, W, 2), dtype=np.float32)
    g[1:-1, 1:-1, 0] = 0.5 * (f[1:-1, 2:] - f[1:-1, :-2])
    g[1:-1, 1:-1, 1] = 0.5 * (f[2:, 1:-1] - f[:-2, 1:-1])
    return g


def pyramid_down(I: np.ndarray) -> np.ndarray:
    """ImagePyramid.h:59-99 — 2x2 box filter in float, truncating cast back to uint8."""
    H, W = I.shape
    Hl, Wl = H // 2, W // 2
    f = I[:2 * Hl, :2 * Wl].astype(np.float32)
    s = np.float32(0.25) * (f[0::2, 0::2] + f[0::2, 1::2] + f[1::2, 0::2] + f[1::2, 1::2])
    return s.astype(np.uint8)


def warp_image(I_ref: np.ndarray, q, t, D: float, fx, fy, cx, cy) -> np.ndarray:
    """generate_synthetic_data.cpp:127-150 — every live pixel warped through the plane Z = D of the keyframe and
    bilinearly sampled (fp32 weights); invalid pixels are 0; result truncated to uint8."""
    H, W = I_ref.shape
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    ray = np.stack([(xx - cx) / fx, (yy - cy) / fy, np.ones_like(xx)], axis=-1)
    R = q_to_R(q)
    m = ray @ R.T
    s = (D - t[2]) / m[..., 2]
    Px = t[0] + s * m[..., 0]
    Py = t[1] + s * m[..., 1]
    Pz = t[2] + s * m[..., 2]
    iz = 1.0 / (Pz + 1e-8)
    u = fx * Px * iz + cx
    v = fy * Py * iz + cy
    valid = (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1)
    u = np.where(valid, u, 0.0)
    v = np.where(valid, v, 0.0)
    xi = u.astype(np.int64)
    yi = v.astype(np.int64)
    dx = (u - xi).astype(np.float32)
    dy = (v - yi).astype(np.float32)
    dxdy = dx * dy
    w00 = np.float32(1.0) - dx - dy + dxdy
    w01 = dx - dxdy
    w10 = dy - dxdy
    w11 = dxdy
    x1 = np.minimum(xi + 1, W - 1)
    y1 = np.minimum(yi + 1, H - 1)
    If = I_ref.astype(np.float32)
    out = w11 * If[y1, x1] + w10 * If[y1, xi] + w01 * If[yi, x1] + w00 * If[yi, xi]
    return np.where(valid, out, 0).astype(np.uint8)


def synthesize_blurred(I_ref, D, fx, fy, cx, cy, k, knots_t, knots_R, t0, dt, cap, exp, num_samples=64) -> np.ndarray:
    """generate_synthetic_data.cpp:152-180 — mean of num_samples warps, float accumulate, round to uint8."""
    acc = np.zeros(I_ref.shape, dtype=np.float32)
    for i in range(num_samples):
        t = cap - 0.5 * exp + i * exp / (num_samples - 1)
        tt, q = spline_pose(k, knots_t, knots_R, t0, dt, t)
        acc += warp_image(I_ref, q, tt, D, fx, fy, cx, cy).astype(np.float32)
    return np.clip(np.rint(acc / np.float32(num_samples)), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------
# problems
# ---------------------------------------------------------------------------------------------------------
def make_gt_spline(n_knots: int, k: int, rng, rot_per_seg=0.02, trans_per_seg=0.375):
    """n knots with small, smooth motion: per segment a rotation <= rot_per_seg rad and a translation <= trans_per_seg."""
    kt = np.zeros((n_knots, 3))
    kR = np.zeros((n_knots, 4))
    kR[:, 3] = 1.0
    base_w = rng.uniform(-1, 1, 3)
    base_w *= rot_per_seg / np.linalg.norm(base_w) * rng.uniform(0.6, 1.0)
    base_v = rng.uniform(-1, 1, 3) * np.array([1.0, 1.0, 0.3])
    base_v *= trans_per_seg / np.linalg.norm(base_v) * rng.uniform(0.6, 1.0)
    kt[0] = rng.uniform(-0.02, 0.02, 3)
    kR[0] = q_exp(rng.uniform(-0.005, 0.005, 3))
    for j in range(1, n_knots):
        wj = base_w * rng.uniform(0.8, 1.2) + rng.uniform(-0.1, 0.1, 3) * rot_per_seg
        vj = base_v * rng.uniform(0.8, 1.2) + rng.uniform(-0.1, 0.1, 3) * trans_per_seg
        kt[j] = kt[j - 1] + vj
        kR[j] = q_mul(kR[j - 1], q_exp(wj))
    return kt, kR


def perturb_knots(kt, kR, rng, sigma=1e-3):
    kt2 = kt + rng.normal(0, sigma, kt.shape)
    kR2 = np.stack([q_mul(q, q_exp(rng.normal(0, sigma, 3))) for q in kR])
    return kt2, kR2


def make_problem(name: str, W: int, H: int, levels: int, P0: int, N: int, n_knots: int, k: int = 2, seed: int = 1234,
                 depth_mode: str = "random", motion_scale: float = 1.0, image: str = "texture", F: int = 1,
                 pattern: Optional[np.ndarray] = None, margin: Optional[int] = None, sigma_init: float = 1e-3,
                 n_gt_samples: int = 64, huber_a: float = 10.0) -> Problem:
    """Build one seeded problem.  The exposure window [t0, t0 + (n_knots - k + 1) * dt] spans all segments of the
    spline, `cap` = middle of the exposure."""
    rng = np.random.default_rng(seed)
    pattern = PATTERN8 if pattern is None else np.asarray(pattern, dtype=np.int32)
    n_seg = n_knots - k + 1
    dt = 1.0 / n_seg
    t_exp = 1.0
    t0 = 0.0
    cap = np.full(F, 0.5 * t_exp)
    exp = np.full(F, t_exp)
    plane_z = 7.5
    fx0 = fy0 = W / 2.0
    cx0, cy0 = W / 2.0, H / 2.0

    gt_t, gt_R = make_gt_spline(n_knots, k, rng, rot_per_seg=0.02 * motion_scale / n_seg,
                                trans_per_seg=0.05 * plane_z * motion_scale / n_seg)
    init_t, init_R = perturb_knots(gt_t, gt_R, rng, sigma_init)

    I0 = make_texture(H, W, seed + 17) if image == "texture" else (shapes_image(H, W) if image == "shapes" else ramp_image(H, W))
    cur0 = [synthesize_blurred(I0, plane_z, fx0, fy0, cx0, cy0, k, gt_t, gt_R, t0, dt, cap[f], exp[f], n_gt_samples)
            for f in range(F)]

    # blur length in level-0 pixels bounds the margin that keeps every sample inside the image
    explicit_margin = margin is not None
    if margin is None:
        flow = 0.0
        for corner in ([0, 0], [W - 1, 0], [0, H - 1], [W - 1, H - 1], [W / 2, H / 2]):
            pts = []
            for tt in np.linspace(0, t_exp * (1 - 1e-9), 9):
                t_, q_ = spline_pose(k, gt_t, gt_R, t0, dt, tt)
                ray = np.array([(corner[0] - cx0) / fx0, (corner[1] - cy0) / fy0, 1.0])
                m = q_to_R(q_) @ ray
                Pp = t_ + (5.0 - t_[2]) / m[2] * m
                pts.append([fx0 * Pp[0] / Pp[2] + cx0 - corner[0], fy0 * Pp[1] / Pp[2] + cy0 - corner[1]])
            flow = max(flow, float(np.abs(np.array(pts)).max()))
        margin = int(np.ceil(flow)) + 8

    lv: List[Level] = []
    I_l, cur_l = I0, cur0
    for l in range(levels):
        if l > 0:
            I_l = pyramid_down(I_l)
            cur_l = [pyramid_down(c) for c in cur_l]
        sc = 2 ** l
        Hl, Wl = H // sc, W // sc
        Pl = max(P0 >> l, 1)
        mg = max(margin // sc, 0) if explicit_margin else max(margin // sc + 4, 5)
        xy = np.stack([rng.uniform(mg, Wl - 1 - mg, Pl), rng.uniform(mg, Hl - 1 - mg, Pl)], axis=1)
        z = rng.uniform(5.0, 10.0, Pl) if depth_mode == "random" else np.full(Pl, plane_z)
        lv.append(Level(H=Hl, W=Wl, fx=fx0 / sc, fy=fy0 / sc, cx=cx0 / sc, cy=cy0 / sc, ref_I=np.ascontiguousarray(I_l),
                        ref_dIxy=image_gradient(I_l), cur_I=[np.ascontiguousarray(c) for c in cur_l],
                        xy=np.ascontiguousarray(xy), z=np.ascontiguousarray(z), pattern=pattern.copy(), N=N))

    seg_start = np.array([int((c - t0) / dt) for c in cap], dtype=np.int32)
    return Problem(name=name, levels=lv, cap=cap, exp=exp, k=k, t0=t0, dt=dt, knots_t=init_t, knots_R=init_R,
                   gt_knots_t=gt_t, gt_knots_R=gt_R, huber_a=huber_a, seg_start=seg_start)


CONFIGS = {
    # BASELINE.json configs[0..4]
    "C1": dict(W=640, H=480, levels=1, P0=2000, N=4, n_knots=2, k=2, seed=1234),
    "C2": dict(W=640, H=480, levels=4, P0=20000, N=16, n_knots=2, k=2, seed=1235, depth_mode="plane"),
    "C3": dict(W=1280, H=720, levels=5, P0=80000, N=32, n_knots=3, k=2, seed=1236),
    "C5": dict(W=640, H=480, levels=1, P0=50000, N=64, n_knots=5, k=2, seed=1238, motion_scale=2.0),
    # extras used by the tests
    "C5cubic": dict(W=640, H=480, levels=1, P0=5000, N=16, n_knots=7, k=4, seed=1239, motion_scale=2.0),
    "tiny": dict(W=160, H=120, levels=2, P0=300, N=8, n_knots=2, k=2, seed=7),
}


def make_config(name: str, **overrides) -> Problem:
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    return make_problem(name, **cfg)
