"""Development probe: globaltimer stamps of every pass of a coarse-to-fine sweep (needs the MBAVO_PROFILE_PHASES build, selected
with MBAVO_LIBRARY).  Persistent form (one sweep_kernel launch) and per-pass form (MBAVO_NO_PERSISTENT=1).
usage: MBAVO_LIBRARY=.../libmbavo_phases.so python scripts/gpu_sweep_timeline.py C3 C2"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402

for name in sys.argv[1:] or ["C3", "C2"]:
    prob = pkg.synth.make_config(name)
    top = len(prob.levels) - 1
    for env in ("", "1"):
        os.environ["MBAVO_NO_PERSISTENT"] = env
        with pkg.Context(limits_for(prob)) as ctx:
            upload_problem(ctx, prob)
            buf = (C.c_ulonglong * (64 * 16))()
            a = (top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4)
            for _ in range(3):
                ctx.gn_sweep(*a, chain=True)
            ctx.lib.mbavo_debug_phase_times(ctx._h, buf)
            ctx.gn_sweep(*a, chain=True)
            ctx.lib.mbavo_debug_phase_times(ctx._h, buf)
            t = np.array(list(buf), dtype=np.float64).reshape(64, 16)
            t0 = t[0, 0]
            print(f"{name} sweep timeline,", "persistent launch" if env == "" else "one launch per pass", "| persistent sweeps:", ctx.persistent_sweeps())
            for r in range(2 * (top + 1)):
                row = t[r]
                us = lambda i: (row[i] - t0) / 1e3  # noqa: E731
                line = (f"  pass {r} ({'H' if r % 2 == 0 else 'C'} level {top - r // 2}): entry +{us(0):7.1f}  records ready +{us(2):7.1f}  loaded +{us(3):7.1f}  "
                        f"batches done(block 0) +{us(6):7.1f}  partials +{us(7):7.1f}  ticket(last) +{us(8):7.1f}  summed +{us(9):7.1f}  end +{us(10):7.1f}")
                if r % 2 == 0 and row[11] > 0:
                    line += f" | built +{(row[11] - row[9]) / 1e3:.1f} solved +{(row[12] - row[9]) / 1e3:.1f} model +{(row[13] - row[9]) / 1e3:.1f} cand +{(row[14] - row[9]) / 1e3:.1f}"
                    if row[15] > 0:
                        line += f" records +{(row[15] - row[9]) / 1e3:.1f}"
                print(line)
