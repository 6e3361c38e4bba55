"""Host mirror of the reference's training entry point (train_bc.py:69-125 `Workspace.run`, utils/py_utils.py:41-79
`Every` / `Timer`, data/robomimic_latent_data.py:116-147 window sampling) around `LDPAgent.update`.

What is here: the update loop with the reference's cadence objects, the latent-sequence window sampler with its edge
padding, a CSV metrics logger, `.npz` snapshots of `agent.get_params()` and the data-only half of `eval`.  What is not:
Hydra configs, environment rollouts (robosuite / ALOHA are absent from this image) and orbax checkpoints - those are
the reference's control plane (out of scope).

Under torchrun every rank runs the same loop on its own shard of each global batch (`batch_size % world == 0`,
train_bc.py:73); `LDPAgent.update` all-reduces the gradients.
"""
from __future__ import annotations

import csv
import json
import time
from collections import defaultdict
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import params as P


class Every:
    """utils/py_utils.py:41-53."""

    def __init__(self, every, action_repeat: int = 1):
        self._every, self._action_repeat = every, action_repeat

    def __call__(self, step: int) -> bool:
        if self._every is None or self._every == -1:
            return False
        return step % (self._every // self._action_repeat) == 0


class Timer:
    """utils/py_utils.py:55-79."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.counts, self.times, self.start_times = defaultdict(int), defaultdict(float), {}

    def tick(self, key):
        if key in self.start_times:
            raise ValueError(f"Timer is already ticking for key: {key}")
        self.start_times[key] = time.time()

    def tock(self, key):
        if key not in self.start_times:
            raise ValueError(f"Timer is not ticking for key: {key}")
        self.counts[key] += 1
        self.times[key] += time.time() - self.start_times.pop(key)

    def get_average_times(self, reset: bool = True):
        ret = {k: self.times[k] / self.counts[k] for k in self.counts}
        if reset:
            self.reset()
        return ret


class LatentSequenceDataset:
    """Window sampler of data/robomimic_latent_data.py: every timestep of every demo is a sample; a sample is the window
    [index - n_frame_stack + 1, index + seq_length) clipped to its demo and padded by repeating the first / last frame
    (:116-147).  `episodes[demo] = {'obs': {key: (L, ...)}, 'actions': (L, A)}` with L = T_ep + 1 rows: the reference
    appends the last next_obs and repeats the last action (:95-110); `process_sdvae_data` latents already have L rows."""

    def __init__(self, episodes: Dict[str, Dict[str, Any]], obs_keys: Sequence[str], seq_length: int, n_frame_stack: int = 1):
        self.obs_keys, self.seq_length, self.n_frame_stack = list(obs_keys), int(seq_length), int(n_frame_stack)
        self.data: Dict[str, np.ndarray] = {}
        starts, lengths, index_to_demo = [], [], []
        pos = 0
        demos = list(episodes)
        for d, name in enumerate(demos):
            n = int(np.asarray(episodes[name]["actions"]).shape[0])
            starts.append(pos)
            lengths.append(n)
            index_to_demo += [d] * n
            pos += n
        for key in ["actions"] + self.obs_keys:
            parts = [np.asarray(episodes[n]["actions"] if key == "actions" else episodes[n]["obs"][key]) for n in demos]
            for n, a, ln in zip(demos, parts, lengths):
                if a.shape[0] != ln:
                    raise ValueError(f"demo {n}: key {key} has {a.shape[0]} rows, actions have {ln}")
            self.data[key] = np.concatenate(parts, axis=0).astype(np.float32) if parts else np.zeros((0,), np.float32)
        self._starts, self._lengths, self._index_to_demo = starts, lengths, index_to_demo
        self.total_n_sequences = pos

    def __len__(self):
        return self.total_n_sequences

    @classmethod
    def from_latent_file(cls, latent_path, episodes: Dict[str, Dict[str, Any]], rgb_keys: Sequence[str],
                         lowdim_keys: Sequence[str], seq_length: int, n_frame_stack: int = 1,
                         data_name: str = "rm_lift") -> "LatentSequenceDataset":
        """The reference's dataset assembly (data/robomimic_latent_data.py:85-110): image observations come from the
        latent file `process_sdvae_data` wrote (`data/<demo>/latent/<rgb_key>`, flattened to h*w*c per frame under the
        key `latent_<rgb_key>`), low-dim observations from the demonstration itself with the last `next_obs` row
        appended (robomimic), actions with the last action repeated - so every key has T_ep + 1 rows."""
        path = Path(latent_path)
        if path.suffix == ".npz":
            z = np.load(path)
            read = lambda demo, key: z[f"data/{demo}/latent/{key}"]
        else:
            import h5py  # type: ignore
            f = h5py.File(path, "r")
            read = lambda demo, key: f["data"][demo]["latent"][key][:]
        out: Dict[str, Dict[str, Any]] = {}
        for demo, ep in episodes.items():
            obs: Dict[str, np.ndarray] = {}
            for k in rgb_keys:
                lat = np.asarray(read(demo, k), dtype=np.float32)
                obs[f"latent_{k}"] = lat.reshape(lat.shape[0], -1)
            actions = np.asarray(ep["actions"], dtype=np.float32)
            if "rm" in data_name:
                for k in lowdim_keys:
                    o = np.asarray(ep["obs"][k], dtype=np.float32)
                    obs[k] = np.concatenate([o, np.asarray(ep["next_obs"][k], dtype=np.float32)[-1:]], axis=0)
                actions = np.concatenate([actions, actions[-1:]], axis=0)
            else:
                for k in lowdim_keys:
                    obs[k] = np.asarray(ep["obs"][k], dtype=np.float32)
            out[demo] = {"obs": obs, "actions": actions}
        return cls(out, [f"latent_{k}" for k in rgb_keys] + list(lowdim_keys), seq_length, n_frame_stack)

    def get_item(self, index: int) -> Dict[str, Any]:
        d = self._index_to_demo[index]
        lo, hi = self._starts[d], self._starts[d] + self._lengths[d]
        s0 = max(index - self.n_frame_stack + 1, lo)
        s1 = min(index + self.seq_length, hi)
        pad0 = max(self.n_frame_stack - (index - s0 + 1), 0)
        pad1 = max(self.seq_length - (s1 - index), 0)

        def window(key):
            seq = self.data[key][s0:s1]
            return np.concatenate([seq[:1]] * pad0 + [seq] + [seq[-1:]] * pad1, axis=0)
        return {"actions": window("actions")[self.n_frame_stack - 1:], "obs": {k: window(k) for k in self.obs_keys}}

    # ---- vectorised path: the whole dataset lives on one device and a batch is one gather per key -------------------
    def to(self, device) -> "LatentSequenceDataset":
        """Keep the flat arrays as tensors on `device` (a latent dataset is small: 8x8x4 floats per frame)."""
        self._dev = {k: torch.as_tensor(v).to(device) for k, v in self.data.items()}
        self._starts_t = torch.as_tensor(np.asarray(self._starts, dtype=np.int64)).to(device)
        self._lengths_t = torch.as_tensor(np.asarray(self._lengths, dtype=np.int64)).to(device)
        self._demo_of_t = torch.as_tensor(np.asarray(self._index_to_demo, dtype=np.int64)).to(device)
        return self

    def gather_batch(self, indices) -> Dict[str, Any]:
        """`get_item` for a vector of indices at once: row r of the window of sample i is the flat row
        clamp(i - (n_frame_stack - 1) + r, demo_start, demo_end - 1) - clamping IS the reference's edge padding
        (repeat the first / last frame of the demo)."""
        if not hasattr(self, "_dev"):
            self.to("cpu")
        idx = torch.as_tensor(indices, dtype=torch.int64, device=self._starts_t.device)
        demo = self._demo_of_t[idx]
        lo = self._starts_t[demo]
        hi = lo + self._lengths_t[demo] - 1
        fs = self.n_frame_stack
        r = torch.arange(fs - 1 + self.seq_length, device=idx.device)
        rows = torch.minimum(torch.maximum(idx[:, None] - (fs - 1) + r[None, :], lo[:, None]), hi[:, None])
        return {"actions": self._dev["actions"][rows[:, fs - 1:]],
                "obs": {k: self._dev[k][rows] for k in self.obs_keys}}

    def sample_batch_fast(self, batch_size: int, rng: np.random.Generator, rank: int = 0, world: int = 1) -> Dict[str, Any]:
        """Same draw and same result as `sample_batch`, gathered on the dataset's device."""
        if batch_size % world:
            raise AssertionError("batch_size % n_devices != 0 (train_bc.py:73)")
        idx = rng.integers(0, self.total_n_sequences, size=batch_size)
        per = batch_size // world
        return self.gather_batch(idx[rank * per:(rank + 1) * per])

    def sample_batch(self, batch_size: int, rng: np.random.Generator, rank: int = 0, world: int = 1) -> Dict[str, Any]:
        """One GLOBAL batch of `batch_size` windows drawn from `rng` (identical on every rank); returns rank's shard."""
        if batch_size % world:
            raise AssertionError("batch_size % n_devices != 0 (train_bc.py:73)")
        idx = rng.integers(0, self.total_n_sequences, size=batch_size)
        per = batch_size // world
        items = [self.get_item(int(i)) for i in idx[rank * per:(rank + 1) * per]]
        return {"actions": torch.from_numpy(np.stack([it["actions"] for it in items])),
                "obs": {k: torch.from_numpy(np.stack([it["obs"][k] for it in items])) for k in self.obs_keys}}


class CSVLogger:
    """Averages the metrics logged between two dumps and appends one row per dump to `<dir>/<ty>.csv`."""

    def __init__(self, log_dir):
        self.dir = Path(log_dir)
        self.dir.mkdir(parents=True, exist_ok=True)
        self._acc: Dict[str, Dict[str, List[float]]] = defaultdict(lambda: defaultdict(list))

    def log_metrics(self, metrics: Dict[str, Any], step: int, ty: str = "train"):
        for k, v in metrics.items():
            if v is None:
                continue
            self._acc[ty][k].append(float(v))

    def dump(self, step: int, ty: str = "train") -> Dict[str, float]:
        row = {"step": step, **{k: float(np.mean(v)) for k, v in self._acc[ty].items()}}
        path = self.dir / f"{ty}.csv"
        old: List[Dict[str, Any]] = []
        if path.exists():
            with open(path, newline="") as f:
                old = list(csv.DictReader(f))
        fields = list(dict.fromkeys([k for r in old for k in r] + list(row)))
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=fields)
            w.writeheader()
            for r in old + [row]:
                w.writerow(r)
        self._acc[ty].clear()
        return row


DEFAULTS = dict(seed=0, batch_size=256, n_grad_steps=500000, log_every_step=10, dump_every_step=200,
                save_every_step=100000, eval_every_step=10000, n_eval_batches=10)   # train_bc.yaml:10-35


class Workspace:
    """train_bc.py `Workspace`: owns the step counter, cadence objects, logger and snapshot directory."""

    def __init__(self, agent, dataset: LatentSequenceDataset, work_dir, eval_dataset: Optional[LatentSequenceDataset] = None,
                 **cfg):
        unknown = set(cfg) - set(DEFAULTS)
        if unknown:
            raise TypeError(f"unknown config keys {sorted(unknown)}")
        self.cfg = {**DEFAULTS, **cfg}
        self.agent, self.dataset, self.eval_dataset = agent, dataset, eval_dataset
        self.work_dir = Path(work_dir)
        self.ckpt_dir = self.work_dir / "ckpt"
        self.ckpt_dir.mkdir(parents=True, exist_ok=True)
        self.logger, self.timer = CSVLogger(self.work_dir), Timer()
        self.step = 0
        import torch.distributed as dist
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0

    def run(self):
        c = self.cfg
        rng = np.random.default_rng(c["seed"])
        every = {k: Every(c[f"{k}_every_step"]) for k in ("eval", "save", "log", "dump")}
        start = time.time()
        metrics: Dict[str, Any] = {}
        while self.step < c["n_grad_steps"]:
            self.timer.tick("time/update_loop")
            sample = self.dataset.sample_batch_fast if hasattr(self.dataset, "_dev") else self.dataset.sample_batch
            batch = sample(c["batch_size"], rng, self.rank, self.world)
            update_seed = int(rng.integers(0, 2 ** 31 - 1))
            self.agent, metrics = self.agent.update(batch, update_seed, self.step)
            self.step += 1
            if every["log"](self.step):
                metrics = {k: (v.item() if torch.is_tensor(v) else v) for k, v in metrics.items()}   # the loop's only sync
                self.timer.tock("time/update_loop")
                metrics.update(self.timer.get_average_times())
                metrics["total_time"] = time.time() - start
                if self.rank == 0:
                    self.logger.log_metrics(metrics, self.step, ty="train")
            else:
                self.timer.tock("time/update_loop")
            if every["save"](self.step) and self.rank == 0:
                self.save_snapshot()
            if every["eval"](self.step):
                self.eval(int(rng.integers(0, 2 ** 31 - 1)))
            if every["dump"](self.step) and self.rank == 0:
                self.logger.dump(self.step, ty="train")
        return metrics

    def eval(self, seed: int) -> Dict[str, float]:
        """The data half of train_bc.py:127-160 (`get_metrics`, action MSE / L1 of `sample_action` and `sample`)."""
        self.timer.tick("time/eval")
        out: Dict[str, float] = {}
        if self.eval_dataset is not None:
            rng = np.random.default_rng(seed)
            rows = []
            for i in range(self.cfg["n_eval_batches"] + 1):
                batch = self.eval_dataset.sample_batch(self.cfg["batch_size"], rng, self.rank, self.world)
                m = dict(self.agent.get_metrics(batch, seed + i))
                pred = self.agent.sample_action(batch, seed + i)
                Hh = pred.shape[1]
                tgt = batch["actions"][:, :Hh].to(pred.device)
                m["action_mse"], m["action_l1"] = ((tgt - pred) ** 2).mean(), (tgt - pred).abs().mean()
                if self.agent.use_planner:
                    oh = self.agent.config["obs_horizon"]
                    full, info = self.agent.sample({"obs": {k: v for k, v in batch["obs"].items()}}, seed + i)
                    Hf = full.shape[1]
                    m["full_action_mse"] = ((batch["actions"][:, oh - 1:oh - 1 + Hf].to(full.device) - full) ** 2).mean()
                    if "plan_mse" in info:
                        m["plan_mse"] = info["plan_mse"]
                rows.append({k: float(v) for k, v in m.items()})
            out = {f"evaldata/{k}": float(np.mean([r[k] for r in rows])) for k in rows[0]}
        self.timer.tock("time/eval")
        if self.rank == 0:
            self.logger.log_metrics(out, self.step, ty="eval")
            self.logger.dump(self.step, ty="eval")
        return out

    # ------------------------------------------------------------------ snapshots (train_bc.py:197-231, as .npz)
    def save_snapshot(self) -> Path:
        params = self.agent.get_params()
        flat = {f"planner_params/{k}": v for k, v in P.unnest(params["planner_params"]).items()}
        flat.update({f"idm_params/{k}": v for k, v in P.unnest(params["idm_params"]).items()})
        flat["cfg"] = np.frombuffer(json.dumps({"workspace": self.cfg, "agent": {k: v for k, v in self.agent.config.items()}},
                                               default=str).encode(), dtype=np.uint8)
        flat["step"] = np.asarray(self.step)
        path = self.ckpt_dir / f"{self.step}.ckpt.npz"
        np.savez(path, **flat)
        return path

    def load_snapshot(self, file, restore_keys: Sequence[str] = ("planner_params", "idm_params")):
        with np.load(file) as z:
            trees = {}
            for key in restore_keys:
                pre = key + "/"
                trees[key] = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
            self.step = int(z["step"])
        self.agent.load_params(trees.get("planner_params") or None, trees.get("idm_params") or None)
        return self.agent


# ---------------------------------------------------------------------------------------------------------------------
# command line (the reference's entry point is Hydra-driven: `python train_bc.py agent=ldp_agent data=cfg/rm_lift/latent_img`)
# ---------------------------------------------------------------------------------------------------------------------
RM_LIFT_LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]      # data/cfg/rm_lift/latent_img.yaml:20-24
RM_LIFT_SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2]}


def synthetic_rm_lift_episodes(n_demos: int = 50, length: int = 120, latent_dim: int = 256, seed: int = 0):
    """Stand-in for a robomimic-lift latent dataset (there is no dataset in this image): same keys and shapes."""
    rs = np.random.default_rng(seed)
    return {f"demo_{i}": {"obs": {"latent_agentview_image": rs.normal(0, 3, (length, latent_dim)).astype(np.float32),
                                  **{k: rs.uniform(-1, 1, (length, RM_LIFT_SHAPES[k][0])).astype(np.float32) for k in RM_LIFT_LOWDIM}},
                          "actions": rs.uniform(-1, 1, (length, 7)).astype(np.float32)} for i in range(n_demos)}


def main(argv=None):
    import argparse
    import os
    import torch.distributed as dist
    from .agent import LDPAgent
    ap = argparse.ArgumentParser(description="LDP training loop on the B200-native kernels (rm_lift latent_img shapes)")
    ap.add_argument("--work-dir", default="exp_local/train_bc")
    ap.add_argument("--steps", type=int, default=200, help="n_grad_steps")
    ap.add_argument("--batch-size", type=int, default=256, help="GLOBAL batch (train_bc.yaml:10); divided over ranks")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--latent-dim", type=int, default=256)
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--save-every", type=int, default=-1)
    a = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    shapes = {**RM_LIFT_SHAPES, "latent_agentview_image": [a.latent_dim]}
    norm = {"obs": {"latent_agentview_image": {"min": np.full(a.latent_dim, -10.0, np.float32), "max": np.full(a.latent_dim, 10.0, np.float32)},
                    **{k: {"min": -np.ones(n[0], np.float32), "max": np.ones(n[0], np.float32)} for k, n in RM_LIFT_SHAPES.items()}},
            "actions": {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}}   # data/cfg/rm_lift/latent_img.yaml:53-61
    agent = LDPAgent.create(a.seed, None, {"ac_dim": 7, "all_shapes": shapes}, rgb_obs=["latent_agentview_image"],
                            lowdim_obs=RM_LIFT_LOWDIM, obs_normalization=norm, vae_feature_dim=a.latent_dim, precision=a.precision,
                            decay_steps=max(a.steps, 1001))
    ds = LatentSequenceDataset(synthetic_rm_lift_episodes(latent_dim=a.latent_dim, seed=a.seed),
                               ["latent_agentview_image"] + RM_LIFT_LOWDIM, seq_length=9).to("cuda")
    ws = Workspace(agent, ds, a.work_dir, seed=a.seed, batch_size=a.batch_size, n_grad_steps=a.steps, log_every_step=a.log_every,
                   dump_every_step=a.log_every, save_every_step=a.save_every, eval_every_step=-1)
    t0 = time.time()
    last = ws.run()
    torch.cuda.synchronize()
    if ws.rank == 0:
        dt = time.time() - t0
        print(json.dumps({"steps": ws.step, "loss": float(last.get("loss", float("nan"))), "seconds": dt,
                          "samples_per_sec": ws.step * a.batch_size / dt, "log": str(ws.work_dir / "train.csv")}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
