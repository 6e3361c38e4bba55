"""Host-side profile of LDPAgent.update (single GPU): where the ~4 ms of enqueue time per step go."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench as BN  # noqa: E402
from latent_diffusion_planning_b200.agent import LDPAgent  # noqa: E402

agent = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": BN.RM_SHAPES}, rgb_obs=["latent_agentview_image"], lowdim_obs=BN.RM_LOWDIM,
                        obs_normalization=BN._rm_norm(np), vae_feature_dim=BN.LATENT, obs_horizon=1, pred_horizon=8, action_horizon=4)
g = torch.Generator().manual_seed(5)
b = 256
tb = {"obs": {"latent_agentview_image": (torch.randn(b, 9, BN.LATENT, generator=g) * 3).cuda()}, "actions": torch.randn(b, 9, 7, generator=g).cuda()}
for k in BN.RM_LOWDIM:
    tb["obs"][k] = (torch.rand(b, 9, BN.RM_SHAPES[k][0], generator=g) * 2 - 1).cuda()
for i in range(6):
    agent.update(tb, i, i)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    agent.update(tb, 6 + i, 6 + i)
pr.disable()
host = (time.perf_counter() - t0) / n * 1e3
e1.record()
torch.cuda.synchronize()
print(f"host enqueue {host:.3f} ms/step (under cProfile), device {e0.elapsed_time(e1) / n:.3f} ms/step")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
# device-only time of the planner + idm loss/grad calls, back to back (no python glue between)
ts_p, ts_i = agent._train["planner"], agent._train["idm"]
