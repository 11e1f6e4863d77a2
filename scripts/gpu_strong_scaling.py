"""BASELINE config 4: the 1280x720 / 5-level / 80k-point / 32-sample / 3-control-pose workload (C3) point-sharded across the
ranks of one node — STRONG scaling (total work fixed), packed H/g/cost all-reduced inside the tracking kernel through the
peer-mapped mailboxes.  One process per GPU:

    python scripts/gpu_strong_scaling.py                                              # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/gpu_strong_scaling.py

Timed on the device (CUDA events on the library's stream, barrier + synchronize on both sides, L2 flushed between steps, max over
ranks): one `mbavo_gn_sweep` (Hessian pass + in-kernel solve + cost pass on each of the 5 levels) per step, and the level-0 Hessian
evaluation alone.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    pkg = ge.load_package()
    from mbavo_b200 import api
    from mbavo_b200.parallel import shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = int(os.environ.get("STEPS", "50")), 5
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prob = pkg.synth.make_config("C3")
    lim = api.limits_for(prob)
    lim.device = local_rank
    ctx = pkg.Context(lim)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_frame_times(prob.cap, prob.exp)
    for l, lv in enumerate(prob.levels):
        lo, hi = shard_bounds(lv.P, rank, world)
        ctx.set_level(l, lv, slice(lo, hi))
    if world > 1:
        handle, _ = ctx.shard_export()
        gathered = [None] * world
        dist.all_gather_object(gathered, handle)
        ctx.shard_connect(world, rank, handles=gathered)
        for l, lv in enumerate(prob.levels):
            ctx.shard_set_global_points(l, lv.P)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    top = len(prob.levels) - 1

    def timed(fn, n):
        total = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            total += e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total / n

    def sweep():
        return ctx.gn_sweep(top, 0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, 1e4)

    def hess0():
        return ctx.evaluate(0, prob.k, prob.t0, prob.dt, prob.knots_t, prob.knots_R, prob.huber_a, True)

    timed(sweep, warmup)
    ms_sweep = timed(sweep, steps)
    timed(hess0, warmup)
    ms_h0 = timed(hess0, steps)
    costs = sweep()[0]
    ps_sweep = sum(lv.P * lv.N for lv in prob.levels) * prob.F * 2  # Hessian pass + cost pass on every level
    ps_h0 = prob.levels[0].P * prob.levels[0].N * prob.F
    if rank == 0:
        peak = 6469.6
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f).get("hbm_gbs", peak))
        except Exception:
            pass
        print(json.dumps({"config": "C4: C3 (1280x720, 5 levels, 80k points, 32 samples, 3 control poses) point-sharded, strong scaling",
                          "n_gpus": world, "steps": steps, "device_sweeps": ctx.device_sweeps(),
                          "gn_sweep_ms": ms_sweep, "gn_sweep_point_samples_per_s": ps_sweep / (ms_sweep * 1e-3),
                          "level0_hessian_eval_ms": ms_h0, "level0_hessian_point_samples_per_s": ps_h0 / (ms_h0 * 1e-3),
                          "level0_hessian_algorithmic_GBps_all_gpus": ps_h0 * 288 / (ms_h0 * 1e-3) / 1e9,
                          "level0_hessian_frac_of_hbm_roofline_per_gpu": ps_h0 * 288 / (ms_h0 * 1e-3) / 1e9 / world / peak,
                          "cost_level0": float(costs[-1, 0])}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
