"""Development probe: 2 ranks in one process (threads, mailboxes by pointer) run a chained sharded sweep; compared with the unsharded one."""
import os, sys, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package()
from mbavo_b200 import api
from mbavo_b200.parallel import shard_bounds
from helpers import run_ranks

prob = pkg.synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "tiny")
top = len(prob.levels) - 1
a = (top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4)
with pkg.Context(api.limits_for(prob)) as ctx:
    api.upload_problem(ctx, prob)
    want = ctx.gn_sweep(*a, chain=True)
    print("unsharded", want[0].tolist(), ctx.persistent_sweeps())
world = 2
ctxs = [pkg.Context(api.limits_for(prob)) for _ in range(world)]
ptrs = []
for r, c in enumerate(ctxs):
    c.set_frame_times(prob.cap, prob.exp)
    for l, lv in enumerate(prob.levels):
        lo, hi = shard_bounds(lv.P, r, world)
        c.set_level(l, lv, slice(lo, hi))
    ptrs.append(c.shard_export()[1])
for r, c in enumerate(ctxs):
    c.shard_connect(world, r, mailbox_ptrs=ptrs)
    for l, lv in enumerate(prob.levels):
        c.shard_set_global_points(l, lv.P)
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for i in range(pre):
    ev = run_ranks([lambda c=c: c.evaluate(i % (top + 1), prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, i % 2 == 0) for c in ctxs])
    print("eval cost", [e[0] for e in ev])
for rep in range(2):
    res = run_ranks([lambda c=c: c.gn_sweep(*a, chain=True) for c in ctxs])
    for r, (costs, kt, kR) in enumerate(res):
        print("rank", r, costs.tolist(), "max knot diff vs unsharded", float(np.abs(kt - want[1]).max()), ctxs[r].persistent_sweeps(), ctxs[r].device_sweeps())
