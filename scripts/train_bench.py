"""Times LDPAgent.update on BASELINE config #4 shapes (train_bc.py agent=ldp_agent data=cfg/rm_lift/latent_img:
batch 256 per GPU, horizon 9, latent 8x8x4 + 9 low-dim = D 265, A 7).  Works under torchrun (NCCL gradient all-reduce).

    python scripts/train_bench.py [--batch 256] [--steps 10] [--warmup 3]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latent_diffusion_planning_b200.agent import LDPAgent  # noqa: E402

LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [256]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workspace", action="store_true", help="time the train_bc.Workspace loop (on-device window sampling + update + logging cadence), wall clock")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    norm = {"obs": {"latent_agentview_image": {"min": np.full(256, -10.0, np.float32), "max": np.full(256, 10.0, np.float32)}},
            "actions": {"clip_min": np.full(7, -1.0, np.float32), "clip_max": np.full(7, 1.0, np.float32)}}
    for k in LOWDIM:
        norm["obs"][k] = {"min": np.full(SHAPES[k][0], -1.0, np.float32), "max": np.full(SHAPES[k][0], 1.0, np.float32)}
    ag = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": SHAPES}, rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM,
                         obs_normalization=norm, vae_feature_dim=256, obs_horizon=1, pred_horizon=8, action_horizon=4)
    if a.workspace:
        import tempfile
        import time
        from latent_diffusion_planning_b200 import train_bc as TB
        rs = np.random.default_rng(0)
        eps = {f"demo_{i}": {"obs": {"latent_agentview_image": rs.normal(0, 3, (120, 256)).astype(np.float32),
                                     **{k: rs.uniform(-1, 1, (120, SHAPES[k][0])).astype(np.float32) for k in LOWDIM}},
                             "actions": rs.uniform(-1, 1, (120, 7)).astype(np.float32)} for i in range(50)}
        ds = TB.LatentSequenceDataset(eps, ["latent_agentview_image"] + LOWDIM, seq_length=9).to("cuda")
        with tempfile.TemporaryDirectory() as d:
            ws = TB.Workspace(ag, ds, d, batch_size=a.batch * world, n_grad_steps=a.warmup, log_every_step=10, dump_every_step=-1,
                              save_every_step=-1, eval_every_step=-1)
            ws.run()
            torch.cuda.synchronize()
            ws.cfg["n_grad_steps"] = a.warmup + a.steps
            t0 = time.time()
            ws.run()
            torch.cuda.synchronize()
            ms = (time.time() - t0) / a.steps * 1e3
        if rank == 0:
            print(json.dumps({"workspace_loop_ms_per_step": ms, "batch_per_gpu": a.batch, "n_gpus": world,
                              "samples_per_sec": a.batch * world / ms * 1e3,
                              "config": "train_bc.Workspace.run: on-device window sampling + LDPAgent.update + metrics every 10 steps, wall clock"}))
        if world > 1:
            dist.destroy_process_group()
        return
    g = torch.Generator().manual_seed(rank)
    B = a.batch
    batch = {"obs": {"latent_agentview_image": (torch.randn(B, 9, 256, generator=g) * 3).cuda()},
             "actions": torch.randn(B, 9, 7, generator=g).cuda()}
    for k in LOWDIM:
        batch["obs"][k] = (torch.rand(B, 9, SHAPES[k][0], generator=g) * 2 - 1).cuda()
    losses = []
    for i in range(a.warmup):
        _, m = ag.update(batch, i, i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        _, m = ag.update(batch, a.warmup + i, a.warmup + i)
        losses.append(m["loss"])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms)
        # executed FLOPs: planner forward 313.0 MF/sample nominal, IDM 3.56 MF/row x 8 rows; fwd + dgrad + wgrad = 3x
        flops = 3.0 * (313.0e6 + 8 * 3.56e6) * B * world
        print(json.dumps({"train_step_ms": ms, "batch_per_gpu": B, "n_gpus": world, "samples_per_sec": B * world / ms * 1e3,
                          "tflops_nominal": flops / ms / 1e9, "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
                          "dtype": "f32", "config": "LDPAgent.update, rm_lift latent_img shapes (D=265, T=8, A=7), planner 69.5 M + IDM params"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
