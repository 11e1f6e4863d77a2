"""Development probe: the cross-process sharded sweep of tests/test_gpu_tracking.py with knobs (pre-evaluations before the sweep)."""
import multiprocessing as mp, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def run_rank(rank, world, conn, pre):
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from mbavo_b200 import api
    from mbavo_b200.parallel import shard_bounds
    prob = pkg.synth.make_config("tiny")
    ctx = pkg.Context(api.limits_for(prob))
    ctx.set_frame_times(prob.cap, prob.exp)
    for l, lv in enumerate(prob.levels):
        lo, hi = shard_bounds(lv.P, rank, world)
        ctx.set_level(l, lv, slice(lo, hi))
    conn.send(ctx.shard_export()[0])
    ctx.shard_connect(world, rank, handles=conn.recv())
    for l, lv in enumerate(prob.levels):
        ctx.shard_set_global_points(l, lv.P)
    top = len(prob.levels) - 1
    a = (prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
    out = []
    for i in range(pre):
        out.append(("eval", ctx.evaluate(i % (top + 1), *a, i % 2 == 0)[0]))
    for rep in range(3):
        costs, kt, kR = ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
        out.append(("sweep", costs.tolist(), ctx.device_sweeps(), ctx.persistent_sweeps(), ctx.lib.mbavo_last_error().decode()))
    conn.send(out)

if __name__ == "__main__":
    for pre in (0, 1, 4):
        ctxm = mp.get_context("spawn")
        pipes, procs = [], []
        for r in range(2):
            a, b = ctxm.Pipe()
            p = ctxm.Process(target=run_rank, args=(r, 2, b, pre)); p.start(); pipes.append(a); procs.append(p)
        hs = [c.recv() for c in pipes]
        for c in pipes: c.send(hs)
        for r, c in enumerate(pipes):
            if c.poll(120):
                for item in c.recv(): print("pre", pre, "rank", r, item)
            else:
                print("pre", pre, "rank", r, "TIMEOUT")
        for p in procs: p.join(10)
