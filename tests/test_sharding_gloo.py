"""World-size-2 test of the point-sharding logic on CPU (gloo): each rank evaluates its contiguous block of host-map
points (with the oracle standing in for the kernel), the shards are summed with ONE all-reduce, and the result must
equal the unsharded evaluation.  This is the host logic of mbavo_b200.parallel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from oracle import oracle as O

    pkg = ge.load_package()
    from mbavo_b200.parallel import reduce_packed, shard_bounds

    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = pkg.synth.make_config("tiny")
    orc = O.OracleLib()
    orc.set_num_threads(1)
    results = []
    for level, lv in enumerate(prob.levels):
        lo, hi = shard_bounds(lv.P, rank, world)
        flags = np.zeros(lv.P, dtype=np.uint8)
        flags[3::11] = 1
        nbad = int(flags.sum())
        shard = pkg.synth.Level(**{**lv.__dict__, "xy": np.ascontiguousarray(lv.xy[lo:hi]), "z": np.ascontiguousarray(lv.z[lo:hi])})
        sub = pkg.synth.Problem(**{**prob.__dict__, "levels": [shard]})
        nbad_local = int(flags[lo:hi].sum())
        c, H, g, _ = orc.evaluate(sub, 0, flags=flags[lo:hi], num_bad=nbad_local)
        # the shard normalised by its own residual count; rescale to the GLOBAL count before the sum
        scale = ((hi - lo) - nbad_local) / (lv.P - nbad)
        local = np.concatenate([[c], g, H.reshape(-1)]) * scale

        def all_reduce_sum(v):
            t = torch.from_numpy(v.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t.numpy()

        total = reduce_packed(local, all_reduce_sum)
        results.append(total)
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), *results)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_all_points():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.load_package()
    from mbavo_b200.parallel import shard_bounds

    for P in (1, 7, 8, 2500, 80000, 80001):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(P, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == P
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharded_sum_equals_full(tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import oracle as O

    pkg = ge.load_package()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    prob = pkg.synth.make_config("tiny")
    orc = O.OracleLib()
    for level, lv in enumerate(prob.levels):
        flags = np.zeros(lv.P, dtype=np.uint8)
        flags[3::11] = 1
        c, H, g, _ = orc.evaluate(prob, level, flags=flags, num_bad=int(flags.sum()))
        full = np.concatenate([[c], g, H.reshape(-1)])
        assert np.abs(got[f"arr_{level}"] - full).max() <= 1e-12 * np.abs(full).max()
