"""2-rank check of the data-parallel update (torchrun, NCCL): every rank trains on its half of a global batch with the
gradient all-reduce, and on the whole batch alone (data_parallel=False); parameters after two steps must agree.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_train_check.py
"""
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
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [16]}


def make(precision):
    norm = {"obs": {"latent_agentview_image": {"min": np.full(16, -10.0, np.float32), "max": np.full(16, 10.0, np.float32)}},
            "actions": {"clip_min": np.full(7, -1.0, np.float32), "clip_max": np.full(7, 1.0, np.float32)}}
    for k in LOWDIM:
        norm["obs"][k] = {"min": np.full(SHAPES[k][0], -1.0, np.float32), "max": np.full(SHAPES[k][0], 1.0, np.float32)}
    return LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=(32, 64), diffusion_step_embed_dim=32),
                           rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=norm, vae_feature_dim=16,
                           vae_block_out_channels=(32,) * 6, precision=precision, lr=1e-3, warmup_steps=2, decay_steps=10)


def main():
    os.environ.setdefault("LDP_BUCKET_MB", "0.05")        # tiny networks: force many gradient buckets so the per-bucket events are exercised
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    out = {}
    for precision in ("fp32", "bf16"):
        dp, full = make(precision), make(precision)
        full.data_parallel = False
        Bg = 8 * world
        for step in range(4):          # steps 3 and 4 replay the captured graph (bf16 path), bucket events included
            g = torch.Generator().manual_seed(step)
            batch = {"obs": {"latent_agentview_image": torch.randn(Bg, 9, 16, generator=g) * 3}, "actions": torch.randn(Bg, 9, 7, generator=g)}
            for k in LOWDIM:
                batch["obs"][k] = torch.rand(Bg, 9, SHAPES[k][0], generator=g) * 2 - 1
            lo, hi = rank * 8, (rank + 1) * 8
            shard = {"obs": {k: v[lo:hi] for k, v in batch["obs"].items()}, "actions": batch["actions"][lo:hi]}
            _, m_dp = dp.update(shard, 7 + step, step)
            _, m_full = full.update(batch, 7 + step, step)
        dl = abs(float(m_dp["loss"]) - float(m_full["loss"]))
        dg = abs(float(m_dp["g_norm"]) - float(m_full["g_norm"]))
        dparam = max(float((dp._train[n].params - full._train[n].params).abs().max()) for n in ("planner", "idm"))
        nb = dp._train["planner"].grad_buckets()
        out[precision] = dict(loss_diff=dl, g_norm_diff=dg, param_diff=dparam, loss=float(m_full["loss"]),
                              planner_buckets=len(nb), buckets_with_event=sum(1 for b in nb if b[2]))
    ok = out["fp32"]["param_diff"] < 2e-5 and out["fp32"]["loss_diff"] < 1e-5 and out["bf16"]["param_diff"] < 2e-3
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"dp_equals_full_batch": bool(flag.item() == 1.0), "world": world, **out}))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
