#!/usr/bin/env python
"""bench.py — throughput of the blur-aware tracking hot path (BASELINE.json metric: point-sample residuals/s).

Workload (config.workload): BASELINE.json configs[2] — synthetic 1280x720, 5-level pyramid, 80k points at level 0
(halved per level), 32 exposure samples, 3 control poses (k = 2, the exposure spans two spline segments): the largest
single-GPU configuration.  With N > 1 GPUs the SAME frame is point-sharded over the ranks — BASELINE.json configs[3], the
strong-scaling curve ("scaling": "strong": total work fixed, 80k / N points per rank at level 0).
One STEP = one Gauss–Newton iteration at every pyramid level, coarse to fine: Hessian-pass evaluation at the current
knots -> damped solve (first LM radius) -> cost-only evaluation at the candidate knots -> the finer level starts from the
candidate when it lowered the cost; i.e. 10 evaluations of the fused kernel and 5 solves, ONE C-ABI call
(mbavo_gn_sweep).  A point-sample = one host-map point at one exposure sub-step (= 8 pixel samples); a step processes
sum_l 2 * P_l * N of them.

Legs (own arm):
  parity    before anything is timed: cost / H / g / first LM step of the timed configuration against the CPU oracle
            (gates of SURVEY.md §8d: cost 1e-5, step 1e-4), and at N > 1 the sharded global result against the unsharded
            evaluation of the same frame on rank 0's GPU.  The run aborts if a gate fails.
  value     inputs resident in HBM, K steps through the C-ABI, device-timed (CUDA events on the library's stream)
  e2e       the same K steps, but every step first re-uploads the frame from PINNED host buffers through the C-ABI — the
            level-0 keyframe and live image (the coarser levels, gradients and texels are built on the GPU) and the host-map
            points of every level — and reads costs and knots back: host<->device copies inside the timed region
  roofline  per-kernel durations of the same evaluations (CUDA events around the tracking kernel, on its stream); at
            N > 1 the rank's kernel is timed WITHOUT the exchange and the exchange is reported separately
  gpu_baseline  (N = 1) the reference's own CUDA kernels (oracle/_ref/libmbavo_refcuda.so: src/ba_tracker/*.cu compiled
            unmodified for sm_100a, points chunked to <= 65 535 as compute_hessian_gradients_cost.cu:309 requires) on the
            same frame and the same GPU, first checked against the CPU reference arithmetic
  cpu_baseline  (N = 1) the reference's arithmetic on the host cores (oracle/_ref, else the oracle port), bounded sample
  extra.C2  (N = 1) the round-1 headline (BASELINE.json configs[1]) as a secondary figure, incl. LM to convergence

`--impl reference` times the CPU implementation alone (rank 0 only) on the same workload, with every host core.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C3": "C3: synthetic 1280x720, 5-level pyramid, 80k points (halved per level), 32 exposure samples, 3 control poses; "
          "with N > 1 GPUs the same frame point-sharded over the ranks (C4, strong scaling)",
    "C2": "C2: synthetic 640x480, 4-level pyramid, 20k points (halved per level), 16 exposure samples, 2 control poses",
}
STEP_TEXT = "1 GN iteration per pyramid level, coarse to fine (H pass + damped solve + cost pass, candidate chained)"
ALGO_BYTES_H = 288.0   # algorithmic bytes per point-sample, Hessian pass (8 px * (4 B + 32 B))
ALGO_BYTES_C = 32.0    # cost-only pass
COST_GATE, DELTA_GATE = 1e-5, 1e-4


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the measurement runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        masks = {}
        for k, v in names.items():
            m = getattr(nv, v, None) or getattr(nv, v.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if m is not None:
                masks[k] = m
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, m in masks.items():
                    if r & m:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def point_samples_per_step(prob):
    return sum(2 * lv.P * lv.N * prob.F for lv in prob.levels)


def config_dict(name, prob, world):
    """The `config` object of the JSON line — identical in both arms (the driver compares them)."""
    return {"workload": WORKLOADS[name], "step": STEP_TEXT, "point_samples_per_step": point_samples_per_step(prob),
            "points_level0": prob.levels[0].P, "levels": len(prob.levels), "exposure_samples": prob.levels[0].N,
            "control_knots": prob.n_knots,
            "l2": "GPU arm: flushed between steps with a 256 MiB memset (not timed); the pyramid is L2-resident within a step",
            "parallelism": f"points of every level sharded over {world} GPUs, one all-reduce of [cost, g, H] per evaluation"
                           if world > 1 else "single GPU"}


# ----------------------------------------------------------------------------------------------------------------
# CPU implementation (reference arm and cpu_baseline)
# ----------------------------------------------------------------------------------------------------------------
def cpu_step(lib, O, prob):
    """The same step on the host: per level H-pass, damped solve, candidate, cost-only pass, chain on decrease."""
    kt, kR = prob.knots_t, prob.knots_R
    for level in reversed(range(len(prob.levels))):
        c, H, g, _ = lib.evaluate(prob, level, kt, kR, want_patch_costs=False)
        step, model = O.trust_region_step(H, g, 1e4)
        ct, cR = O.plus(kt, kR, step)
        c2 = lib.evaluate(prob, level, ct, cR, with_hessian=False, want_patch_costs=False)[0]
        if c2 < c:
            kt, kR = ct, cR


def cpu_lib_all_cores(O):
    """The reference's CPU arithmetic with an explicit thread count (torchrun exports OMP_NUM_THREADS=1)."""
    lib = O.best_cpu_lib()
    lib.set_num_threads(host_cores())
    return lib


def run_reference(args, pkg):
    from oracle import oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lib = cpu_lib_all_cores(O)
    prob = pkg.synth.make_config(args.workload)
    for _ in range(args.warmup):
        cpu_step(lib, O, prob)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(lib, O, prob)
    dt = time.perf_counter() - t
    ps = point_samples_per_step(prob)
    value = ps * args.steps / dt
    line = {"impl": "reference", "metric": "point_sample_residuals_per_s", "value": value, "unit": "point-samples/s",
            "n_gpus": max(args.gpus, world), "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, prob, max(args.gpus, world)),
            "gn_iters_per_s": len(prob.levels) * args.steps / dt,
            "cpu_baseline": {"value": value, "unit": "point-samples/s", "cores": lib.num_threads(), "kind": lib.kind,
                             "sample": f"{args.steps} full steps of the workload, OpenMP over points, {lib.num_threads()} threads"},
            "e2e": {"value": value, "unit": "point-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# GPU baseline: the reference's own CUDA kernels on the same frame
# ----------------------------------------------------------------------------------------------------------------
def run_gpu_baseline(prob, O, reps=5):
    """oracle/_ref/libmbavo_refcuda.so (the reference's src/ba_tracker/*.cu, unmodified, sm_100a) driven as the reference
    drives them (5 launches + 5 device syncs + blocking copies per evaluation; spline_update_step.cpp:97-349), one
    Hessian evaluation + one cost evaluation per level = the work of one bench step.  Checked first against the CPU
    reference arithmetic (same per-frame knot-window semantics).  None when the library was not built."""
    path = os.path.join(ROOT, "oracle", "_ref", "libmbavo_refcuda.so")
    if not os.path.exists(path) or not O.RefLib.available():
        return None
    C = ctypes
    lib = C.CDLL(path)
    lib.mbavo_refcuda_create.restype = C.c_void_p
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    cpu = O.RefLib()
    cpu.set_num_threads(host_cores())
    maxP = max(lv.P for lv in prob.levels)
    maxN = max(lv.N for lv in prob.levels)
    h = lib.mbavo_refcuda_create(1, maxN, maxP, 8, 16, prob.k)
    if not h:
        return {"unavailable": "mbavo_refcuda_create failed (scratch allocation)"}
    h = C.c_void_p(h)
    n = prob.n_knots
    kt, kR = np.ascontiguousarray(prob.knots_t), np.ascontiguousarray(prob.knots_R)
    seg = np.ascontiguousarray(prob.seg_start, dtype=np.int32)
    H, g, cost = np.zeros((6 * n, 6 * n)), np.zeros(6 * n), C.c_double(0)
    out = {"what": "the reference's own CUDA kernels (src/ba_tracker/*.cu unmodified, sm_100a), driven as the reference drives them; "
                   "one Hessian + one cost evaluation per level = the work of one step (wall clock incl. their device syncs and copies)",
           "levels": []}
    total = 0.0
    worst_c = worst_h = 0.0
    try:
        for level, lv in enumerate(prob.levels):
            cur = (C.c_void_p * 1)(lv.cur_I[0].ctypes.data)
            rc = lib.mbavo_refcuda_set_level(h, lv.H, lv.W, C.c_double(lv.fx), C.c_double(lv.fy), C.c_double(lv.cx), C.c_double(lv.cy),
                                             C.c_void_p(lv.ref_I.ctypes.data), C.c_void_p(lv.ref_dIxy.ctypes.data), cur, 1, dp(prob.cap),
                                             dp(prob.exp), dp(lv.xy), dp(lv.z), lv.P, C.c_void_p(lv.pattern.ctypes.data), lv.S, lv.N)
            if rc != 0:
                return {"unavailable": f"mbavo_refcuda_set_level rc={rc}"}

            def ref_eval(with_h):
                return lib.mbavo_refcuda_evaluate(h, C.c_double(prob.t0), C.c_double(prob.dt), dp(kt), dp(kR), n,
                                                  seg.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(prob.huber_a), 0, C.byref(cost),
                                                  dp(H) if with_h else None, dp(g) if with_h else None)

            if ref_eval(True) != 0:
                return {"unavailable": "reference kernels failed"}
            c_cpu, H_cpu, g_cpu, _ = cpu.evaluate(prob, level, want_patch_costs=False)
            err_c = abs(cost.value - c_cpu) / abs(c_cpu)
            err_h = float(np.abs(H - H_cpu).max() / np.abs(H_cpu).max())
            worst_c, worst_h = max(worst_c, err_c), max(worst_h, err_h)
            assert err_c <= 1e-6 and err_h <= 1e-6, f"reference CUDA kernels disagree with the reference CPU arithmetic: {err_c} {err_h}"
            ref_eval(False)
            t = time.perf_counter()
            for _ in range(reps):
                ref_eval(True)
            t_h = (time.perf_counter() - t) / reps
            t = time.perf_counter()
            for _ in range(reps):
                ref_eval(False)
            t_c = (time.perf_counter() - t) / reps
            total += t_h + t_c
            out["levels"].append({"level": level, "points": lv.P, "hessian_ms": t_h * 1e3, "cost_ms": t_c * 1e3})
    finally:
        lib.mbavo_refcuda_destroy(h)
    out["ms_per_step"] = total * 1e3
    out["value"] = point_samples_per_step(prob) / total
    out["unit"] = "point-samples/s"
    out["checked_against_cpu_reference"] = {"cost_rel_max": worst_c, "H_rel_max": worst_h}
    return out


# ----------------------------------------------------------------------------------------------------------------
# own arm
# ----------------------------------------------------------------------------------------------------------------
def shard_slice(P, world, rank):
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


class Arm:
    """One workload on this rank's GPU: context, pinned host copies, the timed loops."""

    def __init__(self, pkg, torch, dist, name, world, rank, local_rank, stream):
        from mbavo_b200.api import Limits

        self.pkg, self.torch, self.dist = pkg, torch, dist
        self._upload = self._sweep = None  # prepared C calls (api.prepare_frame / prepare_gn_sweep)
        self.world, self.rank, self.dev, self.stream = world, rank, torch.device("cuda", local_rank), stream
        self.prob = prob = pkg.synth.make_config(name)   # every rank generates the same frame
        self.ps_step = point_samples_per_step(prob)       # global: the whole frame
        self.slices = [shard_slice(lv.P, world, rank) for lv in prob.levels]
        self.n_levels = len(prob.levels)
        P_mine = max(s.stop - s.start for s in self.slices)
        lim = Limits(max_num_frames=1, max_num_virtual_poses_per_frame=max(lv.N for lv in prob.levels), max_num_keypoints=P_mine,
                     max_patch_size=8, max_num_ctrl_knots=prob.n_knots, device=local_rank)
        self.ctx = ctx = pkg.Context(lim)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_frame_times(prob.cap, prob.exp)

        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        self.pinned = []
        for lv, sl in zip(prob.levels, self.slices):
            self.pinned.append(pkg.synth.Level(H=lv.H, W=lv.W, fx=lv.fx, fy=lv.fy, cx=lv.cx, cy=lv.cy, ref_I=pin(lv.ref_I),
                                               ref_dIxy=lv.ref_dIxy, cur_I=[pin(c) for c in lv.cur_I], xy=pin(lv.xy[sl]),
                                               z=pin(lv.z[sl]), pattern=pin(lv.pattern), N=lv.N))
        # per step: level-0 keyframe + live image, this rank's points / pattern of every level, and per pose-kernel launch the
        # spline state (launch parameter, ~2.3 KB); back: 4 scalars per level + the final knots as (value, tag) pairs
        self.h2d_step = self.pinned[0].ref_I.nbytes + sum(c.nbytes for c in self.pinned[0].cur_I) + \
            sum(lv.xy.nbytes + lv.z.nbytes + lv.pattern.nbytes for lv in self.pinned) + (self.n_levels + 1) * 2304
        self.d2h_step = self.n_levels * 4 * 16 + 7 * prob.n_knots * 16
        self.upload_all()
        if world > 1:
            # one-shot all-reduce inside the tracking kernel: exchange the mailbox IPC handles once, then every C-ABI call is
            # collective (the NCCL-after-the-kernel form is mbavo_b200.parallel.ShardedEvaluator; measured in round 1)
            handle, _ = ctx.shard_export()
            gathered = [None] * world
            dist.all_gather_object(gathered, handle)
            ctx.shard_connect(world, rank, handles=gathered)
            dist.barrier()  # every rank has mapped every mailbox before the first collective call
            for l, lv in enumerate(prob.levels):
                ctx.shard_set_global_points(l, lv.P)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def upload_all(self):
        """The frame from pinned host memory through the C-ABI: mbavo_set_frame (level-0 keyframe + live image on the context's
        stream, pyramid / gradients / texels built on the GPU, points of every level on a second stream), no synchronisation —
        the pinned buffers live as long as this object and the sweep that follows is the blocking call."""
        if self._upload is None:  # (argument marshalling once: the pinned buffers never move)
            self._upload = self.ctx.prepare_frame(self.n_levels, self.pinned[0].ref_I, self.pinned[0].cur_I, self.pinned, async_upload=True)
        self._upload()

    def evaluate(self, level, kt, kR, with_h):
        p = self.prob
        return self.ctx.evaluate(level, p.k, p.t0, p.dt, kt, kR, p.huber_a, with_h)

    def sweep(self):
        p = self.prob
        if self._sweep is None:
            self._sweep = self.ctx.prepare_gn_sweep(self.n_levels - 1, 0, p.k, p.t0, p.dt, p.knots_t, p.knots_R, p.huber_a, 1e4, chain=True)
        costs, kt, kR = self._sweep()
        return costs.copy(), kt.copy(), kR.copy()

    def timed(self, nsteps, with_upload):
        """Device time (ms) of nsteps steps: CUDA events on the stream the kernels run on, L2 flushed (not timed) between
        steps; max over ranks."""
        torch, dist = self.torch, self.dist
        total = 0.0
        self.upload_all(), self.sweep()  # (creates the prepared calls; untimed)
        upload, sweep = self._upload, self._sweep
        launches0 = self.ctx.kernel_launches()
        for _ in range(nsteps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.world > 1:
                dist.barrier()
            torch.cuda.synchronize(self.dev)
            e0.record(self.stream)
            if with_upload:
                upload()
            sweep()  # (blocking: returns when the costs and knots of the sweep have arrived in host memory)
            e1.record(self.stream)
            torch.cuda.synchronize(self.dev)
            total += e0.elapsed_time(e1)
        self.last_launches = self.ctx.kernel_launches() - launches0  # kernels launched inside the timed regions
        if self.world > 1:
            t = torch.tensor([total], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total

    def sweep_kernel_times(self, iters):
        """The persistent sweep kernel of the timed steps, launch by launch: CUDA events around the kernel (ms) and, from the
        globaltimer stamps its passes leave, the duration of every pass (us, [level coarse-first][Hessian, cost]).  None when the
        sweep does not run as one launch (MBAVO_NO_PERSISTENT, windows without an instantiation)."""
        ctx = self.ctx
        if ctx.persistent_sweeps() == 0:
            return None
        ctx.enable_kernel_timing(True)
        ms, passes = [], []
        for it in range(iters + 3):
            self.flush.zero_()
            if self.world > 1:
                self.dist.barrier()
            self.sweep()
            if it >= 3:
                ms.append(ctx.last_kernel_ms())
                passes.append(ctx.sweep_pass_times())
        ctx.enable_kernel_timing(False)
        return sum(ms) / len(ms), np.mean(np.stack(passes), axis=0)

    def kernel_times(self, iters):
        """Per-evaluation tracking-kernel durations (ms) of the step's evaluations, L2 flushed before every step.  Sharded:
        the rank's kernel is launched WITHOUT the exchange (mbavo_evaluate_async), so the figure is a kernel time."""
        ctx, p, torch = self.ctx, self.prob, self.torch
        ctx.enable_kernel_timing(True)
        kms = {}
        scratch = torch.zeros(ctx.packed_len(8), dtype=torch.float64, device=self.dev) if self.world > 1 else None
        for it in range(iters + 3):
            self.flush.zero_()
            for level in reversed(range(self.n_levels)):
                for with_h in (True, False):
                    if self.world > 1:
                        nres = p.levels[level].P * p.F * p.levels[level].S
                        ctx.evaluate_async(level, p.k, p.t0, p.dt, p.knots_t, p.knots_R, p.huber_a, with_h, nres, scratch.data_ptr())
                    else:
                        self.evaluate(level, p.knots_t, p.knots_R, with_h)
                    ms = ctx.last_kernel_ms()
                    if it >= 3:
                        kms.setdefault((level, "H" if with_h else "C"), []).append(ms)
        ctx.enable_kernel_timing(False)
        return {k: sum(v) / len(v) for k, v in kms.items()}

    def close(self):
        self.ctx.close()


def check_parity(arm, O, api):
    """Cost / H / g / first LM step of the timed configuration against the CPU oracle (the C restatement: per-sample knot
    segments), every level; sharded: all ranks hold the same global result, rank 0 also compares it with the unsharded
    evaluation on its own GPU.  Returns the `parity` block; raises when a gate fails."""
    p, world, rank = arm.prob, arm.world, arm.rank
    out = {"gates": {"cost_rel": COST_GATE, "first_lm_step_rel": DELTA_GATE}, "levels": []}
    got = []
    for level in range(arm.n_levels):
        c, H, g = arm.evaluate(level, p.knots_t, p.knots_R, True)     # collective when sharded
        c2 = arm.evaluate(level, p.knots_t, p.knots_R, False)[0]
        got.append((c, H, g, c2))
    sweep_sharded = arm.sweep() if world > 1 else None   # collective: the timed call itself, checked against the unsharded sweep below
    if rank != 0:
        return None
    orc = O.OracleLib()
    orc.set_num_threads(host_cores())
    single = None
    if world > 1:
        single = arm.pkg.Context(api.limits_for(p))
        api.upload_problem(single, p)
    for level, (c, H, g, c2) in enumerate(got):
        c_ref, H_ref, g_ref, _ = orc.evaluate(p, level, want_patch_costs=False)
        d = O.trust_region_step(H.copy(), g, 1e4)[0]
        d_ref = O.trust_region_step(H_ref.copy(), g_ref, 1e4)[0]
        rec = {"level": level, "cost_rel": abs(c - c_ref) / abs(c_ref), "cost_only_rel": abs(c2 - c_ref) / abs(c_ref),
               "first_lm_step_rel": float(np.linalg.norm(d - d_ref) / np.linalg.norm(d_ref)),
               "H_max_rel": float(np.abs(H - H_ref).max() / np.abs(H_ref).max()),
               "g_max_rel": float(np.abs(g - g_ref).max() / np.abs(g_ref).max())}
        if single is not None:
            cs, Hs, gs = single.evaluate(level, p.k, p.t0, p.dt, p.knots_t, p.knots_R, p.huber_a, True)
            rec["sharded_vs_unsharded_cost_rel"] = abs(c - cs) / abs(cs)
            rec["sharded_vs_unsharded_H_max_rel"] = float(np.abs(H - Hs).max() / np.abs(Hs).max())
            # (shard boundaries regroup the fp32 sums over 32-pixel chunks, and a shard small enough to split its exposure samples over
            # the lanes sums them in another order: ~1e-8)
            assert rec["sharded_vs_unsharded_cost_rel"] <= 1e-6 and rec["sharded_vs_unsharded_H_max_rel"] <= 1e-6, rec
        assert rec["cost_rel"] <= COST_GATE and rec["cost_only_rel"] <= COST_GATE, ("cost parity", rec)
        assert rec["first_lm_step_rel"] <= DELTA_GATE, ("first-LM-step parity", rec)
        out["levels"].append(rec)
    if single is not None:
        top = arm.n_levels - 1
        cs, kts, kRs = single.gn_sweep(top, 0, p.k, p.t0, p.dt, p.knots_t, p.knots_R, p.huber_a, 1e4, chain=True)
        out["sharded_sweep_vs_unsharded"] = {"costs_rel_max": float(np.abs(sweep_sharded[0] - cs).max() / np.abs(cs).max()),
                                             "knots_abs_max": float(max(np.abs(sweep_sharded[1] - kts).max(), np.abs(sweep_sharded[2] - kRs).max()))}
        assert out["sharded_sweep_vs_unsharded"]["costs_rel_max"] <= 1e-5 and out["sharded_sweep_vs_unsharded"]["knots_abs_max"] <= 1e-5, out
        single.close()
    out["cost_rel_max"] = max(r["cost_rel"] for r in out["levels"])
    out["first_lm_step_rel_max"] = max(r["first_lm_step_rel"] for r in out["levels"])
    out["oracle"] = "oracle/libmbavo_oracle.so (C restatement, pinned to oracle/_ref and tests/golden)"
    return out


def run_own(args, pkg):
    import torch
    import torch.distributed as dist

    from mbavo_b200 import api
    from oracle import oracle as O   # checker only: parity block and cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one non-default stream for everything: the library's kernels, torch's events / L2 flush
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    arm = Arm(pkg, torch, dist, args.workload, world, rank, local_rank, stream)
    prob, ctx = arm.prob, arm.ctx
    parity = check_parity(arm, O, api)
    if world > 1:
        dist.barrier()

    W = max(args.warmup, 3)
    with ClockSampler(local_rank) as clocks:
        arm.timed(W, False)
        ms_value = arm.timed(args.steps, False)
        launches = arm.last_launches
        arm.timed(W, True)
        ms_e2e = arm.timed(args.steps, True)
        sweep_times = arm.sweep_kernel_times(max(3, min(args.steps, 30)))
        avg = arm.kernel_times(max(3, min(args.steps, 30)))
        exchange = None
        if world > 1:
            # the same level-0 Hessian evaluation WITH the in-kernel exchange (collective): the difference is the exchange + peer wait
            ctx.enable_kernel_timing(True)
            ts = []
            for it in range(13):
                arm.flush.zero_()
                dist.barrier()
                arm.evaluate(0, prob.knots_t, prob.knots_R, True)
                if it >= 3:
                    ts.append(ctx.last_kernel_ms())
            ctx.enable_kernel_timing(False)
            exchange = {"level0_hessian_kernel_with_exchange_ms": sum(ts) / len(ts), "level0_hessian_kernel_alone_ms": avg[(0, "H")],
                        "exchange_and_peer_wait_us": (sum(ts) / len(ts) - avg[(0, "H")]) * 1e3}
    clk = clocks.summary()

    value = arm.ps_step * args.steps / (ms_value * 1e-3)
    e2e_value = arm.ps_step * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = measured_peak_gbs()
    P_mine = [sl.stop - sl.start for sl in arm.slices]
    bytes_of = {(l, t): (ALGO_BYTES_H if t == "H" else ALGO_BYTES_C) * P_mine[l] * prob.levels[l].N * prob.F
                for l in range(arm.n_levels) for t in ("H", "C")}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(f"{args.workload}_sweep_kernel_dram_bytes") if world == 1 else None
    except Exception:
        pass
    note = ("algorithmic bytes = 288 B (Hessian pass) / 32 B (cost pass) per point-sample (SURVEY §8d); the pyramid is L2-resident, "
            "so DRAM traffic is far below it")
    if sweep_times is not None:
        # the dominant kernel of the step IS the step: one persistent launch runs every pass of every level
        k_ms, pass_us = sweep_times
        total_bytes = sum(bytes_of.values())
        achieved = total_bytes / (k_ms * 1e-3) / 1e9
        top = arm.n_levels - 1
        per_pass = {}
        for li in range(arm.n_levels):
            for j, t in enumerate("HC"):
                lvl = top - li
                us = float(pass_us[li, j])
                per_pass[f"L{lvl}{t}"] = {"us": us, "GB/s": bytes_of[(lvl, t)] / (us * 1e-6) / 1e9,
                                          "frac": bytes_of[(lvl, t)] / (us * 1e-6) / 1e9 / peak}
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": f"sweep_kernel<k={prob.k}> (persistent: Hessian pass + solve + cost pass of all {arm.n_levels} levels in one launch)"
                              + (" (this rank's shard, exchanges included)" if world > 1 else ""),
                    "kernel_ms": k_ms, "algorithmic_bytes_per_launch": total_bytes, "peak_source": peak_src,
                    "kernel_share_of_step": k_ms / (ms_value / args.steps),
                    "timed_by": "CUDA events around the launch (serialised behind the pose kernel while events bracket it)",
                    "passes": per_pass,
                    "passes_timed_by": "device globaltimer stamps at the pass releases inside the same launches: a pass lasts from the "
                                       "release of the previous pass to its own (record load, batches, reduction, solve, pose records included)",
                    "dominant_pass": {"pass": "L0H", **per_pass["L0H"], "algorithmic_bytes": bytes_of[(0, "H")]},
                    "standalone_pass_kernels_ms": {f"L{k[0]}{k[1]}": v for k, v in sorted(avg.items())},
                    "standalone_level0_hessian_frac": bytes_of[(0, "H")] / (avg[(0, "H")] * 1e-3) / 1e9 / peak,
                    "note": note + "; `standalone_*` are the same passes as separate track_kernel launches (the non-persistent path), CUDA events"}
    else:
        kernel_ms_per_step = sum(avg.values())
        dom = max(avg, key=lambda k: avg[k])  # dominant kernel launch: level-0 Hessian pass
        achieved = bytes_of[dom] / (avg[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": f"track_kernel<k={prob.k},{'hessian' if dom[1] == 'H' else 'cost'}> level {dom[0]}"
                              + (" (this rank's shard, exchange excluded)" if world > 1 else ""),
                    "kernel_ms": avg[dom], "algorithmic_bytes_per_launch": bytes_of[dom], "peak_source": peak_src,
                    "kernel_ms_per_step_all_launches": kernel_ms_per_step,
                    "kernel_share_of_step": kernel_ms_per_step / (ms_value / args.steps),
                    "per_kernel_ms": {f"L{k[0]}{k[1]}": v for k, v in sorted(avg.items())}, "note": note}
    costs, kt_out, kR_out = arm.sweep()

    line = {"metric": "point_sample_residuals_per_s", "value": value, "unit": "point-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms_value / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, prob, world),
            "gn_iters_per_s": len(prob.levels) * args.steps / (ms_value * 1e-3),
            "e2e": {"value": e2e_value, "unit": "point-samples/s", "h2d_bytes_per_step": int(arm.h2d_step),
                    "d2h_bytes_per_step": int(arm.d2h_step), "ms_per_step": ms_e2e / args.steps,
                    "calls": "per step: mbavo_set_frame(MBAVO_UPLOAD_ASYNC) from pinned host buffers + mbavo_gn_sweep (blocking: costs and "
                             "knots back in host memory), C-ABI through ctypes with the argument structures marshalled once"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "parity": parity,
            "sweep_costs_coarse_to_fine": [[float(a), float(b)] for a, b in costs],
            "accumulation": "fp32 per sample, fp64 patch centres and sums across pixels",
            "roofline": roofline}
    if exchange:
        line["exchange"] = exchange

    if world == 1 and rank == 0 and not args.no_extras:
        line["gpu_baseline"] = run_gpu_baseline(prob, O)
        line["extra"] = {"C2": run_c2_extra(pkg, torch, dist, api, O, local_rank, stream, args)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lib = cpu_lib_all_cores(O)
        cpu_step(lib, O, prob)
        t0, n = time.perf_counter(), 0
        while n < 3 or (time.perf_counter() - t0 < 12.0 and n < 2000):
            cpu_step(lib, O, prob)
            n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": point_samples_per_step(prob) * n / dt, "unit": "point-samples/s",
                                "cores": lib.num_threads(), "kind": lib.kind, "ms_per_step": dt / n * 1e3,
                                "sample": f"{n} full steps of the same workload in {dt:.1f} s (OpenMP over points)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    arm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c2_extra(pkg, torch, dist, api, O, local_rank, stream, args):
    """BASELINE.json configs[1] (the round-1 headline) as a secondary figure: the same step on C2, LM to convergence through
    mbavo_optimize_level, and the new-keyframe set-up (pyramid + semi-dense selection), each next to the host reference."""
    arm = Arm(pkg, torch, dist, "C2", 1, 0, local_rank, stream)
    prob, ctx = arm.prob, arm.ctx
    steps = min(args.steps, 100)
    arm.timed(3, False)
    ms = arm.timed(steps, False)
    arm.timed(3, True)
    ms_e2e = arm.timed(steps, True)
    out = {"workload": WORKLOADS["C2"], "ms_per_step": ms / steps, "value": arm.ps_step * steps / (ms * 1e-3),
           "e2e_ms_per_step": ms_e2e / steps, "e2e_value": arm.ps_step * steps / (ms_e2e * 1e-3), "unit": "point-samples/s"}
    avg = arm.kernel_times(10)
    out["per_kernel_ms"] = {f"L{k[0]}{k[1]}": v for k, v in sorted(avg.items())}
    lv0 = prob.levels[0]
    peak, _ = measured_peak_gbs()
    out["level0_hessian_roofline_frac"] = ALGO_BYTES_H * lv0.P * lv0.N / (avg[(0, "H")] * 1e-3) / 1e9 / peak
    # Gauss-Newton / LM to convergence (BASELINE config 2: "full GN-to-convergence"), wall clock
    arm.upload_all()
    t_lm = time.perf_counter()
    _, _, summ = pkg.optimize_trajectory(ctx, prob)
    t_lm = time.perf_counter() - t_lm
    out["lm_to_convergence"] = {"ms": t_lm * 1e3, "lm_iterations": int(sum(s_["num_iterations"] for s_ in summ)),
                                "evaluations": int(sum(s_["num_evaluations"] for s_ in summ)),
                                "final_cost_level0": float(summ[-1]["final_cost"])}
    arm.close()
    out["shim"] = run_shim_leg(prob)
    depth = np.full((lv0.H, lv0.W), 7.5, dtype=np.float32)
    with pkg.Context(api.limits_for(prob)) as kctx:
        for it in range(13):
            if it == 3:
                t_kf = time.perf_counter()
            kctx.set_keyframe_pyramid(len(prob.levels), lv0.ref_I)
            sel = kctx.select_points(len(prob.levels), depth, lv0.fx, lv0.fy, lv0.cx, lv0.cy, lv0.pattern, lv0.N, 25.0, 30, 30)
        t_kf = (time.perf_counter() - t_kf) / 10
    out["keyframe_setup"] = {"ms": t_kf * 1e3, "points_selected": sel}
    if not args.no_cpu_baseline:
        lib = cpu_lib_all_cores(O)
        t_lm = time.perf_counter()
        _, _, traces = O.optimize_trajectory(lib, prob)
        out["lm_to_convergence"]["cpu_reference_ms"] = (time.perf_counter() - t_lm) * 1e3
        if O.RefSelect.available():
            rs = O.RefSelect()
            t_kf = time.perf_counter()
            for _ in range(5):
                rs.select_points(lv0.ref_I, len(prob.levels), 25.0, 30, 30, depth, max_points=4096)
            out["keyframe_setup"]["cpu_reference_ms"] = (time.perf_counter() - t_kf) / 5 * 1e3
    return out


def run_shim_leg(prob):
    """The drop-in path of INTEGRATION.md §1: the reference's own entry point evaluate_cost_hessian_gradient (same name, same
    arguments, storages poked with raw cudaMemcpy) implemented by mba-vo_b200/host/spline_update_step.cpp on top of the C-ABI,
    called from a C++ program (tests/cpp/dropin_main.cpp) on level 0 of this workload: wall clock per evaluation, with the
    level re-derived on every call (default) and with the opt-in cache.  None when the program cannot be built."""
    import subprocess
    import tempfile

    try:
        binp = os.path.join(ROOT, "tests", "cpp", "dropin_main")
        libdir = os.path.join(ROOT, "mba-vo_b200", "lib")
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++14", os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-I/usr/local/cuda/include",
                        "-L" + libdir, "-L/usr/local/cuda/lib64", "-lmbavo_b200", "-lcudart", "-Wl,-rpath," + libdir,
                        "-Wl,-rpath,/usr/local/cuda/lib64", "-o", binp], check=True, capture_output=True)
        lv = prob.levels[0]
        flags = np.zeros(lv.P, dtype=np.uint8)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "problem.bin")
            with open(path, "wb") as f:
                np.array([lv.H, lv.W, lv.P, lv.S, lv.N, prob.n_knots, prob.k, 0], dtype=np.int32).tofile(f)
                np.array([lv.fx, lv.fy, lv.cx, lv.cy, prob.cap[0], prob.exp[0], prob.t0, prob.dt, prob.huber_a], dtype=np.float64).tofile(f)
                for a in (lv.ref_I, lv.ref_dIxy, lv.cur_I[0], lv.xy, lv.z, lv.pattern, prob.knots_t, prob.knots_R, flags):
                    np.ascontiguousarray(a).tofile(f)
            r = subprocess.run([binp, path], capture_output=True, text=True, timeout=300)
        got = json.loads(r.stdout)
        return {"what": "evaluate_cost_hessian_gradient (reference signature) per evaluation on level 0, wall clock, C++ caller",
                "us_per_evaluation": got["shim_us"]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the gpu_baseline and extra.C2 legs")
    args = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    if args.impl == "reference":
        run_reference(args, pkg)
    else:
        run_own(args, pkg)


if __name__ == "__main__":
    main()
