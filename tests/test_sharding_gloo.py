"""CPU, world_size 2 (gloo): the multi-GPU path's host logic -- contiguous env shards that
start on 256-env blocks, shard-invariant synthetic inputs, all-gather of torques/statistics
and max-over-ranks timing -- with the oracle standing in for the device solve."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import osc_b200 as ob
    from osc_b200 import sharding
    import osc_oracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    spec = ob.load_preset("unitree_go2")
    total = 512
    first, count = sharding.shard_range(total, rank, world)
    inp = ob.synth.make_inputs(spec, count, "go2_standing", first_env=first)
    b = orc.OracleBatch(spec, count, orc.default_settings())
    b.setup(inp)
    o = b.step(inp, n_threads=1)
    tq = sharding.all_gather_rows(torch.from_numpy(o["torque"]), total, world)
    stats = sharding.reduce_stats(dict(solved=int((o["status"] == 1).sum()),
                                       iters=int(o["iters"].sum())), world)
    t = sharding.max_over_ranks(1.0 + rank)
    # the 64-byte CUDA IPC handles of osc_gather_create travel through exchange_bytes
    mine = bytes([(7 * rank + i) % 251 for i in range(64)])
    allh = sharding.exchange_bytes(mine, world)
    assert len(allh) == 64 * world and allh[64 * rank:64 * rank + 64] == mine
    assert allh[64 * (1 - rank):64 * (1 - rank) + 64] == bytes([(7 * (1 - rank) + i) % 251 for i in range(64)])
    if rank == 0:
        q.put((tq.numpy(), stats, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    import osc_b200 as ob
    import osc_oracle as orc
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tq, stats, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = ob.load_preset("unitree_go2")
    inp = ob.synth.make_inputs(spec, 512, "go2_standing")
    b = orc.OracleBatch(spec, 512, orc.default_settings())
    b.setup(inp)
    o = b.step(inp)
    np.testing.assert_array_equal(tq, o["torque"])
    assert stats["solved"] == int((o["status"] == 1).sum())
    assert stats["iters"] == int(o["iters"].sum())
    assert t == 2.0


def test_shard_ranges_cover_and_align():
    from osc_b200 import sharding
    for total in (256, 1024, 16384 * 8, 1000):
        for world in (1, 2, 4, 8):
            if total % 256 and world > 1:
                continue
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f, c), (f2, _) in zip(spans, spans[1:]):
                assert f + c == f2
            assert all(f % 256 == 0 for f, _ in spans)
