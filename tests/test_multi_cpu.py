"""world_size-2 `gloo` tests of the N>1 host logic: landmark sharding + all-reduce of the reduced camera
system (checked with the CPU oracle standing in for the per-rank compute), and the replica layout of the extractor bench."""
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


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from oracle import _ba_bind as B
    from corb_slam_b200.synth import ba_problem, ba_shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.lib()
    prob = ba_problem(15, 900, seed=4, n_fusion=6)
    shard = ba_shard(prob, rank, world)
    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}

    def allreduce(arr, op):
        t = torch.from_numpy(arr)  # shares memory with the oracle's buffer
        dist.all_reduce(t, op=ops[op])

    out, info = B.solve(shard, 6, allreduce=allreduce)
    # every rank must hold the same poses; gather the points back in landmark order
    poses = [torch.zeros(out["pose_t"].shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(poses, torch.from_numpy(np.ascontiguousarray(out["pose_t"])))
    same = all(torch.equal(poses[0], p) for p in poses)
    q.put((rank, same, info["trial_accepted"], out["pose_t"], shard["_point_ids"], out["point_xyz"]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ba_over_gloo_matches_single_rank():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    from oracle import _ba_bind as B
    from corb_slam_b200.synth import ba_problem
    oracle.lib()
    prob = ba_problem(15, 900, seed=4, n_fusion=6)
    full, info = B.solve(prob, 6)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=180) for _ in range(2)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    pts = np.zeros_like(full["point_xyz"])
    for rank, same, accepted, pose_t, ids, xyz in got:
        assert same, "ranks disagree on the poses after the all-reduced solve"
        assert accepted == info["trial_accepted"]
        np.testing.assert_allclose(pose_t, full["pose_t"], atol=1e-8)
        pts[ids] = xyz
    np.testing.assert_allclose(pts, full["point_xyz"], atol=1e-8)


def test_shard_partition_is_exact():
    sys.path.insert(0, ROOT)
    from corb_slam_b200.synth import ba_problem, ba_shard
    prob = ba_problem(12, 500, seed=1, n_fusion=4)
    for world in (2, 3, 8):
        shards = [ba_shard(prob, r, world) for r in range(world)]
        ids = np.concatenate([s["_point_ids"] for s in shards])
        assert sorted(ids.tolist()) == list(range(500))                       # every landmark on exactly one rank
        assert sum(len(s["edge_pose"]) for s in shards) == len(prob["edge_pose"])  # with all of its edges
        for s in shards:
            assert len(s["pose_t"]) == 12                                         # poses are replicated
            glob = s["_point_ids"][s["edge_point"]]
            assert (glob % world == (s["_point_ids"][0] % world if len(s["_point_ids"]) else 0)).all()
