"""Small target for `ncu --set full`: a few Hessian-pass and cost-only evaluations of one config / level.
usage: python scripts/ncu_target.py C3 0 [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MBAVO_NO_GRAPHS"] = "1"  # plain launches: ncu attributes kernels launched from graphs less readably
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
level = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prob = pkg.synth.make_config(name, levels=level + 1)
with pkg.Context(limits_for(prob)) as ctx:
    upload_problem(ctx, prob)
    for _ in range(reps):
        c, H, g = ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
        c2, _, _ = ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
    print(name, level, c, c2)
