"""The TMA-staged cost pass (csrc/track_cost_tma.cu, mbavo_debug_cost_tma) against the product cost pass on the same level:
result, kernel time (CUDA events, median of 10 after warm-up), share of samples served from the tiles, for several box sizes.
usage: python scripts/gpu_tma_variant.py [config:level ...]     (run under gpurun; appends JSON lines to gpurun_out/tma_variant.jsonl)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402

BOXES = [(32, 16), (32, 32), (48, 32), (64, 32), (64, 48), (96, 48)]
only = os.environ.get("TMA_BOX")
if only:
    BOXES = [tuple(int(v) for v in only.split("x"))]
reps = int(os.environ.get("TMA_REPS", "12"))
out = open(os.path.join(ROOT, "gpurun_out", "tma_variant.jsonl"), "a")
cache = {}
for tgt in sys.argv[1:] or ["C3:0", "C5:0", "C3:2", "C2:0"]:
    name, level = tgt.split(":")
    level = int(level)
    if (name, level) not in cache:
        cache[(name, level)] = pkg.synth.make_config(name, levels=level + 1)
    prob = cache[(name, level)]
    lv = prob.levels[level]
    a = (level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
    with pkg.Context(limits_for(prob)) as ctx:
        upload_problem(ctx, prob)
        want = ctx.evaluate(*a, False)[0]
        ctx.enable_kernel_timing(True)
        ms = []
        for _ in range(reps):
            ctx.evaluate(*a, False)
            ms.append(ctx.last_kernel_ms())
        ctx.enable_kernel_timing(False)
        base_us = float(np.median(ms[2:])) * 1e3
        rec = dict(config=name, level=level, P=lv.P, N=lv.N, variant="product (quad-texel gather through L1)", us=base_us, cost=want)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        for bw, bh in BOXES:
            try:
                ts, frac, c = [], 0.0, 0.0
                for _ in range(reps):
                    c, t, frac = ctx.cost_tma(*a, bw, bh)
                    ts.append(t)
            except pkg.MbavoError as e:
                print(f"{tgt} box {bw}x{bh}: {e}", flush=True)
                continue
            rec = dict(config=name, level=level, P=lv.P, N=lv.N, variant=f"TMA tiles {bw}x{bh}", us=float(np.median(ts[2:])) * 1e3,
                       tile_fraction=frac, cost=c, cost_rel_vs_product=abs(c - want) / abs(want), vs_product=float(np.median(ts[2:])) * 1e3 / base_us)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
out.close()
