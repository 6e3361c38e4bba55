"""torchrun: kernel timeline (CUPTI via torch.profiler) of a few data-parallel LDPAgent.update steps on rank 0 -> gpurun_out/dp_trace_<mode>.json (reduced)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench as BN  # noqa: E402
from latent_diffusion_planning_b200.agent import LDPAgent  # noqa: E402

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
step = 0
for mode in ("bucketed", "single"):
    agent.bucketed_allreduce = mode == "bucketed"
    for _ in range(6):
        agent.update(tb, step, step); step += 1
    torch.cuda.synchronize(); dist.barrier()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            agent.update(tb, step, step); step += 1
        torch.cuda.synchronize()
    if rank == 0:
        ev = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                ev.append((e.time_range.start, e.time_range.end, e.name[:60], getattr(e, "device_index", 0)))
        ev.sort()
        t0 = ev[0][0]
        out = [(round(s - t0, 1), round(en - s, 1), n) for s, en, n, _ in ev]
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"dp_trace_{mode}.json").write_text(json.dumps(out))
        nccl = [(s, d) for s, d, n in out if "nccl" in n.lower()]
        print(mode, "kernels", len(out), "span us", out[-1][0] + out[-1][1], "nccl kernels", len(nccl), "nccl busy us", sum(d for _, d in nccl))
dist.destroy_process_group()
