"""Development probe: globaltimer stamps of the tracking kernel's phases (needs the MBAVO_PROFILE_PHASES build of the library,
selected with MBAVO_LIBRARY).  usage: MBAVO_LIBRARY=.../libmbavo_phases.so python scripts/gpu_phases.py C1:0 C2:3 C2:0"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402

NAMES = ["entry", "before griddep", "after griddep", "prologue done", "phase A done (first chunk)", "phase B done (first chunk)",
         "batches done", "block partials stored", "last block: ticket won", "last block: partials summed", "last block: published"]
for tgt in sys.argv[1:] or ["C1:0", "C2:3", "C2:0"]:
    name, level = tgt.split(":")
    level = int(level)
    prob = pkg.synth.make_config(name, levels=level + 1)
    with pkg.Context(limits_for(prob)) as ctx:
        upload_problem(ctx, prob)
        buf = (C.c_ulonglong * (64 * 16))()
        ctx.lib.mbavo_debug_phase_times(ctx._h, buf)  # allocates the stamp buffer
        for with_h in (True, False):
            for _ in range(3):
                ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, with_h)
            ctx.lib.mbavo_debug_phase_times(ctx._h, buf)
            t = np.array(list(buf)[:11], dtype=np.float64)
            print(f"{tgt} {'H' if with_h else 'C'}: " + ", ".join(f"{n} +{(t[i] - t[0]) / 1e3:.1f}us" for i, n in enumerate(NAMES) if i > 0))

# timeline of one whole sweep (C2): every tracking kernel's stamps relative to the first kernel's entry
prob = pkg.synth.make_config("C2")
for env in ("", "1"):
    os.environ["MBAVO_NO_DEVICE_SWEEP"] = env
    with pkg.Context(limits_for(prob)) as ctx:
        upload_problem(ctx, prob)
        buf = (C.c_ulonglong * (64 * 16))()
        for _ in range(3):
            ctx.gn_sweep(3, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4)
        ctx.lib.mbavo_debug_phase_times(ctx._h, buf)
        ctx.gn_sweep(3, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4)
        ctx.lib.mbavo_debug_phase_times(ctx._h, buf)
        t = np.array(list(buf), dtype=np.float64).reshape(64, 16)
        t0 = t[0, 0]
        print("sweep timeline, device-resident" if env == "" else "sweep timeline, evaluation by evaluation", "sweeps on device:", ctx.device_sweeps())
        for r in range(8):
            row = t[r]
            print(f"  kernel {r}: entry +{(row[0] - t0) / 1e3:7.1f}us  griddep done +{(row[2] - t0) / 1e3:7.1f}  batches done +{(row[6] - t0) / 1e3:7.1f}  "
                  f"ticket +{(row[8] - t0) / 1e3:7.1f}  summed +{(row[9] - t0) / 1e3:7.1f}  end +{(row[10] - t0) / 1e3:7.1f}"
                  + (f"  | solve: system built +{(row[11] - row[9]) / 1e3:.1f}  factored+solved +{(row[12] - row[9]) / 1e3:.1f}  "
                     f"model +{(row[13] - row[9]) / 1e3:.1f}  candidate +{(row[14] - row[9]) / 1e3:.1f}" if row[11] > row[9] > 0 else ""))
