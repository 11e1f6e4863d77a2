"""GPU end-to-end test of the per-frame driver (mbavo_track_frame, the mirror of BlurAwareDirectTracker::trackFrame,
blur_aware_direct_tracker.cpp:88-203): a synthetic sequence of motion-blurred frames of a textured plane, rendered by
mbavo_synthesize_blurred along a constant-twist trajectory, is tracked with the points the GPU selected on the keyframe; the
recovered poses are compared with the trajectory that rendered the frames, before and after a keyframe change."""
import os
import subprocess

import numpy as np
import pytest

from helpers import run_ranks

pytestmark = pytest.mark.gpu

H, W, LEVELS, D = 240, 320, 3, 7.5
FX = FY = 160.0
CX, CY = 160.0, 120.0
DT_FRAME, EXPOSURE = 0.1, 0.06
TWIST = np.array([1.6, -0.9, 0.5, 0.05, -0.12, 0.2])  # per second: ~5 px of flow per frame, ~3 px of blur per exposure


def rot_angle(qa, qb):
    d = abs(float(np.dot(qa / np.linalg.norm(qa), qb / np.linalg.norm(qb))))
    return 2.0 * np.arccos(min(1.0, d))


def test_track_blurred_sequence(pkg, O, synth):
    from mbavo_b200 import api

    key = synth.make_texture(H, W, seed=21)
    pattern = synth.make_config("C1").levels[0].pattern

    def gt(t):  # T_cur2key(t) = Exp(t * twist)
        return np.concatenate(O.se3_exp(TWIST * t))

    def render(cap, n=32, exposure=EXPOSURE):
        ts = [cap] if n == 1 else [cap - 0.5 * exposure + j * exposure / (n - 1) for j in range(n)]
        return api.synthesize_blurred(key, D, FX, FY, CX, CY, np.array([gt(t) for t in ts]))

    lim = api.Limits(max_num_keypoints=8192, max_num_virtual_poses_per_frame=16, max_patch_size=len(pattern))
    with pkg.Context(lim) as ctx:
        ctx.set_keyframe_pyramid(LEVELS, key)
        counts = ctx.select_points(LEVELS, np.full((H, W), D, np.float32), FX, FY, CX, CY, pattern, 16, 4.0, 6, 6)
        assert counts[0] > 800 and counts[LEVELS - 1] > 80
        trk = api.FrameTracker(ctx, LEVELS, DT_FRAME, 0.0, huber_a=10.0, max_chi_square_error=3.0)
        errs = []
        for i in range(1, 4):
            cap = i * DT_FRAME
            res = trk.track(render(cap), cap, EXPOSURE)
            want = gt(cap)
            errs.append((np.linalg.norm(res["t_cur2key"] - want[:3]), rot_angle(res["q_cur2key"], want[3:])))
            assert res["levels_run"] == (1 << LEVELS) - 1
            assert all(lv["final_cost"] <= lv["initial_cost"] for lv in res["levels"])
            # the isKeyframe statistics grow with the distance from the keyframe; the blur-kernel length stays that of the exposure
            assert res["avg_flow"] > 3.0 * i and 1.0 < res["avg_kernel_len"] < 8.0
            assert np.allclose(res["t_cur2world"], res["t_cur2key"]) and np.allclose(res["q_cur2world"], res["q_cur2key"])
        print("tracking errors (translation, rotation):", errs)
        # after the first frame the velocity is known and the prediction starts close: the twist per second is recovered
        assert np.abs(trk.velocity - TWIST).max() <= 0.05 * np.abs(TWIST).max()
        for e_t, e_r in errs:
            assert e_t <= 1e-3 * D and e_r <= 5e-4, errs  # observed: 1.1e-3 .. 3.3e-3 and 0.6e-4 .. 1.9e-4

        # keyframe change at frame 3 (tracker.cpp:186-199): sharp frame + its depth of the same plane, points re-selected
        cap3 = 3 * DT_FRAME
        trk.new_keyframe(cap3)
        T3 = gt(cap3)
        R3 = O.q_to_R(T3[3:])
        ys, xs = np.mgrid[0:H, 0:W]
        rays = np.stack([(xs - CX) / FX, (ys - CY) / FY, np.ones((H, W))], axis=-1)
        depth3 = ((D - T3[2]) / (rays @ R3[2])).astype(np.float32)  # e_z . (R z r + t) = D
        ctx.set_keyframe_pyramid(LEVELS, render(cap3, n=1))
        counts = ctx.select_points(LEVELS, depth3, FX, FY, CX, CY, pattern, 16, 4.0, 6, 6)
        assert counts[0] > 800
        for i in range(4, 6):
            cap = i * DT_FRAME
            res = trk.track(render(cap), cap, EXPOSURE)
            want = gt(cap)
            e_t, e_r = np.linalg.norm(res["t_cur2world"] - want[:3]), rot_angle(res["q_cur2world"], want[3:])
            print("after the keyframe change:", e_t, e_r, res["avg_flow"])
            assert e_t <= 2e-3 * D and e_r <= 1.5e-3, (e_t, e_r)  # observed: 5.4e-3, 5.6e-4
            assert res["avg_flow"] < 3.0 * (i - 3) + 8.0  # flow is measured from the NEW keyframe


def test_track_blurred_sequence_point_sharded(pkg, O, synth):
    """The same driver on a point-sharded tracker: two ranks (contexts on this GPU, one thread each) select the same points,
    keep their contiguous block, and track collectively — identical poses on both ranks, equal to the single-context run up to
    the summation order."""
    from mbavo_b200 import api

    key = synth.make_texture(H, W, seed=21)
    pattern = synth.make_config("C1").levels[0].pattern
    depth = np.full((H, W), D, np.float32)

    def gt(t):
        return np.concatenate(O.se3_exp(TWIST * t))

    frames = []
    for i in (1, 2):
        cap = i * DT_FRAME
        ts = [cap - 0.5 * EXPOSURE + j * EXPOSURE / 31 for j in range(32)]
        frames.append((cap, api.synthesize_blurred(key, D, FX, FY, CX, CY, np.array([gt(t) for t in ts]))))
    lim = api.Limits(max_num_keypoints=8192, max_num_virtual_poses_per_frame=16, max_patch_size=len(pattern))

    def run(ctx):
        counts = ctx.select_points(LEVELS, depth, FX, FY, CX, CY, pattern, 16, 4.0, 6, 6)
        trk = api.FrameTracker(ctx, LEVELS, DT_FRAME, 0.0, huber_a=10.0, max_chi_square_error=3.0)
        return counts, [trk.track(img, cap, EXPOSURE) for cap, img in frames], [ctx.get_points(l)[0].shape[0] for l in range(LEVELS)]

    with pkg.Context(lim) as solo:
        solo.set_keyframe_pyramid(LEVELS, key)
        counts1, res1, held1 = run(solo)
    assert held1 == counts1
    ctxs = [pkg.Context(lim) for _ in range(2)]
    try:
        for c in ctxs:
            c.set_keyframe_pyramid(LEVELS, key)
        ptrs = [c.shard_export()[1] for c in ctxs]
        for r, c in enumerate(ctxs):
            c.shard_connect(2, r, mailbox_ptrs=ptrs)
        out = run_ranks([lambda c=c: run(c) for c in ctxs])
    finally:
        for c in ctxs:
            c.close()
    (ca, ra, ha), (cb, rb, hb) = out
    assert ca == counts1 and cb == counts1 and [x + y for x, y in zip(ha, hb)] == counts1  # the blocks partition the selection
    for f in range(2):
        for key_ in ("t_cur2key", "q_cur2key"):
            assert np.array_equal(ra[f][key_], rb[f][key_])
            assert np.abs(ra[f][key_] - res1[f][key_]).max() <= 1e-5
        assert ra[f]["avg_flow"] == rb[f]["avg_flow"] and abs(ra[f]["avg_flow"] - res1[f]["avg_flow"]) <= 1e-4 * res1[f]["avg_flow"]
        assert [lv["decisions"] for lv in ra[f]["levels"]] == [lv["decisions"] for lv in res1[f]["levels"]]


def test_plain_c_tracker_loop(tmp_path):
    """examples/track_sequence.c — keyframe set-up, point selection, mbavo_track_frame per frame, the keyframe decision and a
    keyframe change, all through include/mbavo.h from plain C: builds with gcc -std=c99 and tracks its sequence within its own
    tolerances (exit code 0)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "mba-vo_b200", "lib")
    exe = os.path.join(str(tmp_path), "track_sequence")
    subprocess.run(["/usr/bin/gcc", "-O2", "-std=c99", os.path.join(root, "examples", "track_sequence.c"), "-I" + os.path.join(root, "include"),
                    "-L" + libdir, "-lmbavo_b200", "-lm", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "-> new keyframe" in r.stdout and r.stdout.count("frame ") >= 6


@pytest.mark.parametrize("world", [2])
def test_cross_process_sharding(pkg, synth, world):
    """The cross-PROCESS form of point sharding — what torchrun + bench.py use at N > 1: every rank is its own process with its
    own CUDA context, the mailboxes are mapped through CUDA IPC handles (cudaIpcGetMemHandle / cudaIpcOpenMemHandle), no barrier
    between connect and the first collective.  One GPU per rank when the box has several; on a one-GPU box the ranks share GPU 0
    (their kernels time-slice; the library then keeps each sweep pass in its own launch).  Evaluations, a chained sweep and a
    whole collective LM loop give, on every rank, the result of the unsharded context."""
    import multiprocessing as mp

    import shard_worker
    from mbavo_b200 import api

    config = "tiny"
    prob = synth.make_config(config)
    top = len(prob.levels) - 1
    a = (prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
    with pkg.Context(api.limits_for(prob)) as ctx:
        api.upload_problem(ctx, prob)
        want_eval = [ctx.evaluate(l, *a, True) for l in range(top + 1)]
        want_sweep = ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        want_lm = ctx.optimize_level(top, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, huber_a=prob.huber_a)

    mpc = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(world):
        parent, child = mpc.Pipe()
        p = mpc.Process(target=shard_worker.run_rank, args=(r, world, child, config))
        p.start()
        pipes.append(parent)
        procs.append(p)
    try:
        handles = []
        for r, c in enumerate(pipes):
            assert c.poll(120), f"rank {r} did not export its mailbox"
            kind, payload = c.recv()
            assert kind == "handle", payload
            handles.append(payload)
        for c in pipes:
            c.send(handles)
        results = []
        for r, c in enumerate(pipes):
            assert c.poll(180), f"rank {r} did not finish"
            kind, payload = c.recv()
            assert kind == "result", payload
            results.append(payload)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for res in results:
        for l in range(top + 1):
            c, H, g = res["eval"][l]
            cw, Hw, gw = want_eval[l]
            # (a shard boundary that is not a multiple of the 4-point warp batch regroups the fp32 chunk sums: ~1e-8)
            assert abs(c - cw) <= 1e-9 * cw and np.abs(H - Hw).max() <= 1e-6 * np.abs(Hw).max() and np.abs(g - gw).max() <= 1e-6 * np.abs(gw).max()
            assert abs(res["cost_only"][l] - cw) <= 1e-6 * cw
            # bit-identical on every rank: the slots are summed in rank order everywhere
            assert c == results[0]["eval"][l][0] and np.array_equal(H, results[0]["eval"][l][1])
        # the ~1e-8 regrouping difference of H goes through the solve: candidate knots and their costs agree to ~1e-7
        costs, kt, kR = res["sweep"]
        assert np.abs(costs - want_sweep[0]).max() <= 1e-5 * np.abs(want_sweep[0]).max()
        assert np.abs(kt - want_sweep[1]).max() <= 1e-4 and np.abs(kR - want_sweep[2]).max() <= 1e-4  # (300 points: cond(H) ~ 1e6)
        assert np.array_equal(costs, results[0]["sweep"][0]) and np.array_equal(kt, results[0]["sweep"][1])  # identical on every rank
        lkt, lkR, decisions, final_cost, nbad = res["lm"]
        assert decisions == want_lm[2]["decisions"] and nbad == want_lm[2]["num_bad_keypoints"]
        assert np.abs(lkt - want_lm[0]).max() <= 1e-4 and abs(final_cost - want_lm[2]["final_cost"]) <= 1e-5 * want_lm[2]["final_cost"]
    if len({res["device"] for res in results}) == world:
        assert all(res["persistent_sweeps"] == 1 for res in results)  # a GPU per rank: the sweep ran as one launch on each
    else:
        assert all(res["persistent_sweeps"] == 0 for res in results)
