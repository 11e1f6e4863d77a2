"""Worker of tests/test_gpu_tracking.py::test_cross_process_sharding: ONE rank of a point-sharded group in its OWN process
(spawned), connected to the other ranks through CUDA IPC mailbox handles — the path torchrun + bench.py take at N > 1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_rank(rank, world, conn, config):
    """conn: multiprocessing Pipe end to the parent (handle out, all handles in, results out)."""
    try:
        sys.path.insert(0, ROOT)
        import ctypes

        import __graft_entry__ as ge

        pkg = ge.load_package()
        from mbavo_b200 import api
        from mbavo_b200.parallel import shard_bounds

        cudart = ctypes.CDLL("libcudart.so.12")
        ndev = ctypes.c_int(0)
        cudart.cudaGetDeviceCount(ctypes.byref(ndev))
        dev = rank % max(ndev.value, 1)  # one GPU per rank when the box has them, else the ranks share GPU 0 (IPC still applies)
        prob = pkg.synth.make_config(config)
        lim = api.limits_for(prob)
        lim.device = dev
        ctx = pkg.Context(lim)
        ctx.set_frame_times(prob.cap, prob.exp)
        for l, lv in enumerate(prob.levels):
            lo, hi = shard_bounds(lv.P, rank, world)
            ctx.set_level(l, lv, slice(lo, hi))
        handle, _ = ctx.shard_export()
        conn.send(("handle", handle))
        handles = conn.recv()
        ctx.shard_connect(world, rank, handles=handles)  # no barrier needed: mailboxes were zeroed at export
        for l, lv in enumerate(prob.levels):
            ctx.shard_set_global_points(l, lv.P)
        top = len(prob.levels) - 1
        a = (prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
        out = {"device": dev}
        out["eval"] = [ctx.evaluate(l, *a, True) for l in range(top + 1)]
        out["cost_only"] = [ctx.evaluate(l, *a, False)[0] for l in range(top + 1)]
        out["sweep"] = ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        out["persistent_sweeps"] = ctx.persistent_sweeps()
        kt, kR, summ = ctx.optimize_level(top, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, huber_a=prob.huber_a)
        out["lm"] = (kt, kR, summ["decisions"], summ["final_cost"], summ["num_bad_keypoints"])
        ctx.shard_disconnect()
        ctx.close()
        conn.send(("result", out))
    except BaseException as e:  # noqa: BLE001
        import traceback

        conn.send(("error", f"rank {rank}: {e}\n{traceback.format_exc()}"))
