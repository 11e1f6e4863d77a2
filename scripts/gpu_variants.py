"""Kernel-tuning probe (run under gpurun): device time of the Hessian-pass and cost-only tracking kernel per config /
level for the variant selected by the environment (MBAVO_PHASES, MBAVO_NO_TEXELS, MBAVO_LIBRARY), plus parity of the
variant against the oracle on the small configs.  Appends one JSON line per (variant, config, level) to
gpurun_out/variants.jsonl.   usage: python scripts/gpu_variants.py TAG [config:level ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402

PEAK = 6469.6e9
_cache = {}


def problem(name, levels):
    key = (name, levels)
    if key not in _cache:
        _cache[key] = pkg.synth.make_config(name, levels=levels)
    return _cache[key]


def main():
    tag = sys.argv[1]
    targets = sys.argv[2:] or ["C2:0", "C3:0", "C5:0"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "a")
    for tgt in targets:
        name, level = tgt.split(":")
        level = int(level)
        prob = problem(name, level + 1)
        lv = prob.levels[level]
        with pkg.Context(limits_for(prob)) as ctx:
            upload_problem(ctx, prob)
            args = (level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
            c, H, g = ctx.evaluate(*args, True)
            c2, _, _ = ctx.evaluate(*args, False)
            ctx.enable_kernel_timing(True)
            ms_h, ms_c = [], []
            for _ in range(12):
                ctx.evaluate(*args, True)
                ms_h.append(ctx.last_kernel_ms())
                ctx.evaluate(*args, False)
                ms_c.append(ctx.last_kernel_ms())
            ctx.enable_kernel_timing(False)
            ps = lv.P * lv.N * prob.F
            mh, mc = float(np.median(ms_h[2:])), float(np.median(ms_c[2:]))
            rec = dict(tag=tag, config=name, level=level, P=lv.P, N=lv.N, texels=ctx.level_uses_texels(level),
                       us_h=mh * 1e3, us_c=mc * 1e3, us_h_min=min(ms_h) * 1e3, us_c_min=min(ms_c) * 1e3,
                       frac_h=ps * 288 / (mh * 1e-3) / PEAK, frac_c=ps * 32 / (mc * 1e-3) / PEAK,
                       gps_h=ps / (mh * 1e-3), cost=c, cost_only=c2, H00=float(H[0, 0]), gsum=float(np.abs(g).sum()))
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
