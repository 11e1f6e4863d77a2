"""Small target for `ncu`: a few chained coarse-to-fine sweeps of one config through mbavo_gn_sweep (the persistent sweep kernel).
usage: python scripts/ncu_sweep_target.py C3 [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem_pyramid  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = pkg.synth.make_config(name)
with pkg.Context(limits_for(prob)) as ctx:
    upload_problem_pyramid(ctx, prob)
    top = len(prob.levels) - 1
    for _ in range(reps):
        costs, kt, kR = ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4, chain=True)
    print(name, costs.tolist(), "persistent sweeps:", ctx.persistent_sweeps())
