"""torchrun probe: LDPAgent.update at 256 per rank with (a) no exchange, (b) one all-reduce per network after its backward,
(c) bucketed all-reduce started by per-bucket events.  Prints ms per step for each."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench as BN  # noqa: E402
from latent_diffusion_planning_b200.agent import LDPAgent  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    agent = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": BN.RM_SHAPES}, rgb_obs=["latent_agentview_image"], lowdim_obs=BN.RM_LOWDIM,
                            obs_normalization=BN._rm_norm(np), vae_feature_dim=BN.LATENT, obs_horizon=1, pred_horizon=8, action_horizon=4)
    g = torch.Generator().manual_seed(5 + rank)
    b = 256
    tb = {"obs": {"latent_agentview_image": (torch.randn(b, 9, BN.LATENT, generator=g) * 3).cuda()}, "actions": torch.randn(b, 9, 7, generator=g).cuda()}
    for k in BN.RM_LOWDIM:
        tb["obs"][k] = (torch.rand(b, 9, BN.RM_SHAPES[k][0], generator=g) * 2 - 1).cuda()
    step = [0]
    out = {}

    def run(n):
        for _ in range(n):
            agent.update(tb, step[0], step[0])
            step[0] += 1

    def timed(label, n=20):
        run(4)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        t0 = time.perf_counter()
        e0.record(); run(n); e1.record()
        host_ms = (time.perf_counter() - t0) / n * 1e3          # host time to ENQUEUE a step (no sync inside)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[label + "_host_enqueue_ms"] = host_ms
        return float(ms)
    agent.data_parallel = False
    out["no_exchange_ms"] = timed("off")
    agent.data_parallel = True
    agent.bucketed_allreduce = False
    out["single_allreduce_ms"] = timed("single")
    agent.bucketed_allreduce = True
    out["bucketed_ms"] = timed("bucketed")
    out["buckets"] = agent._train["planner"].grad_buckets()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
