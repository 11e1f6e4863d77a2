"""GPU baseline of BASELINE.md §3.2 (run under gpurun): the REFERENCE's own CUDA kernels (oracle/_ref/libmbavo_refcuda.so,
compiled unmodified from /root/reference for sm_100a behind a restated orchestration, every wrapper followed by the
reference's cudaDeviceSynchronize) on the same inputs and the same B200 as the product library: first validated against
the oracle, then timed (wall clock per evaluation, as the reference runs).  Writes gpurun_out/ref_cuda_baseline.json."""
import ctypes as C
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

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmbavo_refcuda.so"))
lib.mbavo_refcuda_create.restype = C.c_void_p
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    prob = pkg.synth.make_config(name)
    orc = O.OracleLib()
    out = {"config": name, "levels": []}
    maxP = max(lv.P for lv in prob.levels)
    h = C.c_void_p(lib.mbavo_refcuda_create(1, 64, maxP, 8, 16, prob.k))
    with pkg.Context(limits_for(prob)) as ctx:
        upload_problem(ctx, prob)
        for level, lv in enumerate(prob.levels):
            cur = (C.c_void_p * 1)(lv.cur_I[0].ctypes.data)
            rc = lib.mbavo_refcuda_set_level(h, lv.H, lv.W, C.c_double(lv.fx), C.c_double(lv.fy), C.c_double(lv.cx), C.c_double(lv.cy),
                                             C.c_void_p(lv.ref_I.ctypes.data), C.c_void_p(lv.ref_dIxy.ctypes.data), cur, 1, dp(prob.cap),
                                             dp(prob.exp), dp(lv.xy), dp(lv.z), lv.P, C.c_void_p(lv.pattern.ctypes.data), lv.S, lv.N)
            assert rc == 0, rc
            n = prob.n_knots
            H = np.zeros((6 * n, 6 * n))
            g = np.zeros(6 * n)
            cost = C.c_double(0)
            seg = np.zeros(1, dtype=np.int32)
            kt, kR = np.ascontiguousarray(prob.knots_t), np.ascontiguousarray(prob.knots_R)

            def ref_eval(with_h):
                return lib.mbavo_refcuda_evaluate(h, C.c_double(prob.t0), C.c_double(prob.dt), dp(kt), dp(kR), n,
                                                  seg.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(prob.huber_a), 0, C.byref(cost),
                                                  dp(H) if with_h else None, dp(g) if with_h else None)

            assert ref_eval(True) == 0
            c_ref, H_ref, g_ref, _ = orc.evaluate(prob, level)
            err_c = abs(cost.value - c_ref) / c_ref
            err_H = float(np.abs(H - H_ref).max() / np.abs(H_ref).max())
            reps = 5
            t = time.perf_counter()
            for _ in range(reps):
                ref_eval(True)
            t_h = (time.perf_counter() - t) / reps
            t = time.perf_counter()
            for _ in range(reps):
                ref_eval(False)
            t_c = (time.perf_counter() - t) / reps
            a = (level, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a)
            for _ in range(3):
                ctx.evaluate(*a, True)
                ctx.evaluate(*a, False)
            t = time.perf_counter()
            for _ in range(50):
                ctx.evaluate(*a, True)
            o_h = (time.perf_counter() - t) / 50
            t = time.perf_counter()
            for _ in range(50):
                ctx.evaluate(*a, False)
            o_c = (time.perf_counter() - t) / 50
            rec = dict(level=level, P=lv.P, N=lv.N, ref_vs_oracle_cost_rel=err_c, ref_vs_oracle_H_rel=err_H,
                       ref_cuda_hessian_ms=t_h * 1e3, ref_cuda_cost_ms=t_c * 1e3, ours_hessian_ms=o_h * 1e3, ours_cost_ms=o_c * 1e3,
                       speedup_hessian=t_h / o_h, speedup_cost=t_c / o_c)
            print(json.dumps(rec), flush=True)
            out["levels"].append(rec)
    lib.mbavo_refcuda_destroy(h)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ref_cuda_baseline_{name}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
