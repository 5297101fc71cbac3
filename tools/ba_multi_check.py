"""Multi-GPU check of the landmark-sharded global BA (run under torchrun, one rank per GPU):
every rank solves its shard with the NCCL all-reduce hook; rank 0 also solves the unsharded problem on one GPU and
compares (identical LM accept/reject sequence, poses/points within 1e-7)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from corb_slam_b200 import Optimizer, torch_allreduce
    from corb_slam_b200.synth import ba_problem, ba_shard
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 20000)
    prob = ba_problem(P, L, seed=7)
    shard = ba_shard(prob, rank, world)
    cb = torch_allreduce()
    out, info = Optimizer.BundleAdjustment(shard, 10, bRobust=False, device=local, allreduce=cb)
    pts = torch.zeros((L, 3), dtype=torch.float64, device="cuda")
    pts[torch.from_numpy(shard["_point_ids"]).cuda()] = torch.from_numpy(out["point_xyz"]).cuda()
    dist.all_reduce(pts)
    poses = torch.from_numpy(out["pose_t"]).cuda()
    pmax = poses.clone(); pmin = poses.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX); dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        full, finfo = Optimizer.BundleAdjustment(prob, 10, bRobust=False, device=local)
        ok = info["trial_accepted"] == finfo["trial_accepted"]
        dp = float(np.abs(out["pose_t"] - full["pose_t"]).max())
        dx = float(np.abs(pts.cpu().numpy() - full["point_xyz"]).max())
        spread = float((pmax - pmin).abs().max())
        print("BA multi-GPU check: world=%d accept_equal=%s d_pose=%.3e d_point=%.3e rank_spread=%.3e ms_sharded=%.1f ms_single=%.1f"
              % (world, ok, dp, dx, spread, info["ms_total"], finfo["ms_total"]))
        assert ok and dp < 1e-7 and dx < 1e-6 and spread == 0.0
        print("PASS")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
