"""Target for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel of the library on small inputs —
pack, pyramid, pose, tracking (Hessian + cost, texel and direct-gather variants, ragged shapes, borders, k = 4, multi-segment),
outlier and keyframe statistics, point selection, the per-frame driver, and a two-rank sharded evaluation on one device.   usage: python scripts/sanitize_target.py [--quick]"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200 import api  # noqa: E402
from mbavo_b200.parallel import shard_bounds  # noqa: E402

synth = pkg.synth


def run(prob, pyramid=False):
    with pkg.Context(api.limits_for(prob)) as ctx:
        (api.upload_problem_pyramid if pyramid else api.upload_problem)(ctx, prob)
        for level in range(len(prob.levels)):
            a = (level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
            c, H, g = ctx.evaluate(*a, True)
            c2, _, _ = ctx.evaluate(*a, False)
            ctx.detect_outliers(level, 3.0)
            ctx.evaluate(*a, True)
        top = len(prob.levels) - 1
        costs, kt, kR = ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        cap, exp = float(prob.cap[0]), float(prob.exp[0])
        poses = np.array([np.concatenate(synth.spline_pose(prob.k, prob.gt_knots_t, prob.gt_knots_R, prob.t0, prob.dt, t))
                          for t in (cap, cap - 0.4 * exp, cap + 0.4 * exp)])
        kf = ctx.keyframe_stats(0, poses)
        print(prob.name, "ok", c, c2, "sweeps on device", ctx.device_sweeps(), "persistent", ctx.persistent_sweeps(), "kf", kf, flush=True)
        # round 2: the one-call frame upload (two streams, asynchronous, the persistent sweep waits for the per-level ready flags),
        # the debug dump of the product kernels, the TMA-staged cost pass
        if pyramid:
            for _ in range(2):
                ctx.set_frame(len(prob.levels), prob.levels[0].ref_I, prob.levels[0].cur_I, prob.levels, async_upload=True)
                ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
            ctx.set_frame(len(prob.levels), None, prob.levels[0].cur_I, prob.levels, async_upload=False)
            ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
        lv = prob.levels[0]
        if prob.k == 2 and prob.n_knots <= 3 and lv.S == 8:
            d = ctx.debug_dump(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, prob.F, lv.N, lv.P, lv.S)
            print(prob.name, "debug dump ok", d["cost"], flush=True)
            if lv.W % 16 == 0 and ctx.level_uses_texels(0) == 1:
                print(prob.name, "tma ok", ctx.cost_tma(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 32, 16), flush=True)


probs = [synth.make_config("tiny"),
         synth.make_problem("ragged", W=192, H=144, levels=1, P0=77, N=5, n_knots=2, k=2, seed=3, margin=14,
                            pattern=np.array([[0, 0], [1, 2], [-3, 1], [2, -2], [0, 3]], dtype=np.int32)),
         synth.make_problem("border", W=160, H=120, levels=1, P0=600, N=8, n_knots=2, k=2, seed=77, margin=0, motion_scale=2.0),
         synth.make_problem("cubic", W=160, H=120, levels=1, P0=200, N=8, n_knots=7, k=4, seed=9, margin=16, motion_scale=2.0),
         synth.make_problem("multi", W=160, H=120, levels=2, P0=300, N=16, n_knots=5, k=2, seed=11, margin=16)]
if "--quick" in sys.argv:
    # the persistent sweep (2- and 3-knot windows: odd and even lengths of the packed vector), the asynchronous frame upload, the
    # last block's sum / solve / pose records — a few seconds per sanitizer tool
    run(probs[0])
    run(synth.make_problem("pyr16", W=160, H=128, levels=3, P0=300, N=8, n_knots=3, k=2, seed=6, margin=20), pyramid=True)
    sys.exit(0)
for p in probs:
    run(p)
run(synth.make_problem("pyr", W=162, H=122, levels=3, P0=300, N=4, n_knots=2, k=2, seed=5, margin=20), pyramid=True)
run(synth.make_problem("pyr16", W=160, H=128, levels=3, P0=300, N=8, n_knots=3, k=2, seed=6, margin=20), pyramid=True)
os.environ["MBAVO_NO_PERSISTENT"] = "1"
run(probs[0])
del os.environ["MBAVO_NO_PERSISTENT"]
os.environ["MBAVO_NO_TEXELS"] = "1"
run(probs[0])
del os.environ["MBAVO_NO_TEXELS"]
os.environ["MBAVO_PHASES"] = "4"
run(probs[0])
del os.environ["MBAVO_PHASES"]

img = api.synthesize_blurred(probs[0].levels[0].ref_I, 7.5, 48.0, 48.0, 48.0, 32.0,
                             np.array([[0.01 * i, 0.0, 0.0, 0.0, 0.0, 0.001 * i, 1.0] for i in range(6)]))
print("synth ok", int(img.sum()), flush=True)

# semi-dense point selection (odd image sizes: partial last cells, odd pyramid sizes) and the per-frame driver on the selection
key = synth.make_texture(122, 162, seed=4)
pat = probs[0].levels[0].pattern
with pkg.Context(api.Limits(max_num_keypoints=4096, max_num_virtual_poses_per_frame=8, max_patch_size=len(pat))) as ctx:
    ctx.set_keyframe_pyramid(3, key)
    depth = np.full(key.shape, 7.5, np.float32)
    depth[::7, ::5] = 0.0
    for thr, ch, cw in ((4.0, 6, 6), (0.5, 17, 23), (25.0, 30, 30), (300.0, 6, 6)):
        counts = ctx.select_points(3, depth, 81.0, 81.0, 81.0, 61.0, pat, 8, thr, ch, cw)
        print("select ok", thr, counts, [ctx.get_points(l)[0].shape[0] for l in range(3)], flush=True)
    ctx.select_points(3, depth, 81.0, 81.0, 81.0, 61.0, pat, 8, 2.0, 5, 5)
    trk = api.FrameTracker(ctx, 3, 0.1, 0.0, huber_a=10.0, max_chi_square_error=3.0)
    poses = np.array([[0.02 * i, -0.01 * i, 0.0, 0.0, 0.0, 0.0005 * i, 1.0] for i in range(8)])
    res = trk.track(api.synthesize_blurred(key, 7.5, 81.0, 81.0, 81.0, 61.0, poses), 0.1, 0.05)
    print("track ok", res["t_cur2key"], res["levels_run"], flush=True)

# two ranks on one device
prob = probs[0]
ctxs = [pkg.Context(api.limits_for(prob)) for _ in range(2)]
ptrs = []
for r, ctx in enumerate(ctxs):
    ctx.set_frame_times(prob.cap, prob.exp)
    lo, hi = shard_bounds(prob.levels[0].P, r, 2)
    ctx.set_level(0, prob.levels[0], slice(lo, hi))
    ptrs.append(ctx.shard_export()[1])
for r, ctx in enumerate(ctxs):
    ctx.shard_connect(2, r, mailbox_ptrs=ptrs)
    ctx.shard_set_global_points(0, prob.levels[0].P)
res = [None, None]


def body(i):
    a = (0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
    res[i] = ctxs[i].evaluate(*a, True)[0]
    ctxs[i].detect_outliers(0, 3.0)


th = [threading.Thread(target=body, args=(i,)) for i in range(2)]
[t.start() for t in th]
[t.join() for t in th]
print("sharded ok", res, flush=True)
for c in ctxs:
    c.close()
