"""LDPAgent.update / act at odd batch sizes, switching shapes between calls (graph / workspace rebuild paths): diagnostics."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latent_diffusion_planning_b200.agent import LDPAgent  # noqa: E402

LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [256]}
norm = {"obs": {"latent_agentview_image": {"min": np.full(256, -10.0, np.float32), "max": np.full(256, 10.0, np.float32)}},
        "actions": {"clip_min": np.full(7, -1.0, np.float32), "clip_max": np.full(7, 1.0, np.float32)}}
for k in LOWDIM:
    norm["obs"][k] = {"min": np.full(SHAPES[k][0], -1.0, np.float32), "max": np.full(SHAPES[k][0], 1.0, np.float32)}
ag = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": SHAPES}, rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM,
                     obs_normalization=norm, vae_feature_dim=256, precision="bf16", lr=1e-4, warmup_steps=2, decay_steps=100)
ok = True
step = 0
for B in (256, 7, 256, 1, 100, 255, 256, 256, 256):
    g = torch.Generator().manual_seed(B)
    batch = {"obs": {"latent_agentview_image": torch.randn(B, 9, 256, generator=g) * 3}, "actions": torch.rand(B, 9, 7, generator=g) * 2 - 1}
    for k in LOWDIM:
        batch["obs"][k] = torch.rand(B, 9, SHAPES[k][0], generator=g) * 2 - 1
    try:
        _, m = ag.update(batch, step, step)
        loss = float(m["loss"])
        fin = np.isfinite(loss)
        print(f"update B={B}: loss {loss:.4f} finite={fin}")
        ok &= bool(fin)
    except Exception as e:  # noqa: BLE001
        ok = False
        print(f"update B={B}: FAILED {type(e).__name__}: {str(e)[:300]}")
    step += 1
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
