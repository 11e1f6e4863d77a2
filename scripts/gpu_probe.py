"""Development probe (run under gpurun): parity of the CUDA path against the oracle on the BASELINE configs and
kernel timings.  Not part of the test-suite; writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
from mbavo_b200.api import limits_for, upload_problem  # noqa: E402
from oracle import oracle as O  # noqa: E402


def delta(H, g):
    return O.solve_normal_equation(H * (1 + np.eye(H.shape[0]) * 1e-4), g)


def main():
    names = sys.argv[1:] or ["tiny", "C1", "C2", "C3", "C5", "C5cubic"]
    lib = O.OracleLib()  # per-sample segment semantics (oracle/_ref keeps the reference's per-frame attribution)
    out = {"cpu_lib": lib.kind, "cpu_threads": lib.num_threads(), "configs": {}}
    for name in names:
        t = time.time()
        prob = pkg.synth.make_config(name)
        print(f"[{name}] generated in {time.time() - t:.1f}s: levels={len(prob.levels)} P0={prob.levels[0].P} "
              f"N={prob.levels[0].N} n={prob.n_knots} k={prob.k}", flush=True)
        res = {}
        with pkg.Context(limits_for(prob)) as ctx:
            upload_problem(ctx, prob)
            for level, lv in enumerate(prob.levels):
                t = time.time()
                c_ref, H_ref, g_ref, pc_ref = lib.evaluate(prob, level)
                t_cpu = time.time() - t
                c, H, g = ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
                pc = ctx.patch_costs(level, prob.F, lv.P)
                c2, _, _ = ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
                d_ref, d = delta(H_ref, g_ref), delta(H, g)
                r = dict(
                    cost_rel=abs(c - c_ref) / abs(c_ref), cost_only_rel=abs(c2 - c_ref) / abs(c_ref),
                    H_rel=float(np.abs(H - H_ref).max() / np.abs(H_ref).max()),
                    g_rel=float(np.abs(g - g_ref).max() / np.abs(g_ref).max()),
                    delta_rel=float(np.linalg.norm(d - d_ref) / np.linalg.norm(d_ref)),
                    patch_cost_abs=float(np.abs(pc - pc_ref).max()), cond=float(np.linalg.cond(H_ref)), cpu_s=t_cpu)
                # timings
                ctx.enable_kernel_timing(True)
                ms_h, ms_c = [], []
                for _ in range(5):
                    ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
                    ms_h.append(ctx.last_kernel_ms())
                    ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
                    ms_c.append(ctx.last_kernel_ms())
                ctx.enable_kernel_timing(False)
                t = time.time()
                reps = 20
                for _ in range(reps):
                    ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)
                wall_h = (time.time() - t) / reps * 1e3
                t = time.time()
                for _ in range(reps):
                    ctx.evaluate(level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, False)
                wall_c = (time.time() - t) / reps * 1e3
                ps = lv.P * lv.N * prob.F
                r.update(kernel_ms_h=min(ms_h), kernel_ms_c=min(ms_c), wall_ms_h=wall_h, wall_ms_c=wall_c,
                         point_samples=ps, gps_h=ps / (min(ms_h) * 1e-3), gps_c=ps / (min(ms_c) * 1e-3),
                         roofline_frac_h=ps / (min(ms_h) * 1e-3) * 288 / 6550.7e9)
                print(f"  L{level} P={lv.P}: cost_rel {r['cost_rel']:.2e} delta_rel {r['delta_rel']:.2e} H_rel {r['H_rel']:.2e} "
                      f"g_rel {r['g_rel']:.2e} pc {r['patch_cost_abs']:.2e} | kernel H {min(ms_h) * 1e3:.1f}us C {min(ms_c) * 1e3:.1f}us "
                      f"wall H {wall_h * 1e3:.0f}us C {wall_c * 1e3:.0f}us | {r['gps_h']:.3e} pt-samples/s "
                      f"({r['roofline_frac_h'] * 100:.1f}% algo-HBM) | cpu {t_cpu:.2f}s", flush=True)
                res[f"L{level}"] = r
            if name in ("tiny", "C2"):
                t = time.time()
                kt, kR, summ = pkg.optimize_trajectory(ctx, prob)
                t_gpu = time.time() - t
                t = time.time()
                kt_o, kR_o, traces = O.optimize_trajectory(lib, prob)
                t_cpu = time.time() - t
                res["lm"] = dict(gpu_s=t_gpu, cpu_s=t_cpu, decisions_gpu=[s["decisions"] for s in summ],
                                 decisions_cpu=["".join(tr.decisions) for tr in traces],
                                 final_cost_gpu=[s["final_cost"] for s in summ], final_cost_cpu=[tr.costs[-1] for tr in traces],
                                 knots_t_err=float(np.abs(kt - kt_o).max()), knots_R_err=float(np.abs(kR - kR_o).max()),
                                 gt_t_err=float(np.abs(kt - prob.gt_knots_t).max()))
                print("  LM:", json.dumps(res["lm"]), flush=True)
        out["configs"][name] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
