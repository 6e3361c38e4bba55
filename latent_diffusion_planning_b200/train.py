"""Host side of the training step (scope row N1): the reference's `TrainState` + `optax.adam` pair for one network
(agent/ldp_agent.py:580-600), on flat float32 device buffers the CUDA library reads and updates in place.

The reference differentiates `loss` with `jax.grad` under GSPMD (`train_bc.py:73`: the global batch is sharded over
devices and `jnp.mean` runs over all of it).  Here every rank holds an equal shard, `ldp_*_loss_grad` adds the shard's
gradients into the flat gradient buffer, `allreduce_grads` sums that one buffer over ranks (NCCL on GPUs, gloo in the
CPU tests) and the 1/world factor is folded into the Adam kernel's `grad_scale`.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _native as N
from . import params as P


def warmup_cosine_decay_schedule(init_value: float, peak_value: float, warmup_steps: int, decay_steps: int,
                                 end_value: float = 0.0) -> Callable[[int], float]:
    """optax.warmup_cosine_decay_schedule as the reference builds it (agent/ldp_agent.py:580-587: init = end_lr,
    peak = lr): linear warm-up joined at `warmup_steps` with a cosine decay over `decay_steps - warmup_steps`."""
    def schedule(count: int) -> float:
        if count < warmup_steps:
            frac = 1.0 - max(count, 0) / warmup_steps
            return (init_value - peak_value) * frac + peak_value
        n = max(decay_steps - warmup_steps, 1)
        c = min(count - warmup_steps, n)
        alpha = 0.0 if peak_value == 0.0 else end_value / peak_value
        return peak_value * ((1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * c / n)) + alpha)
    return schedule


def allreduce_grads(grads: torch.Tensor, group=None) -> float:
    """Sum the flat gradient buffer over ranks; returns the factor (1/world) that turns the sum of per-shard mean
    gradients into the gradient of the global-batch mean.  No-op (factor 1) without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


def unflatten_params(spec, blob: np.ndarray) -> Dict[str, np.ndarray]:
    out, pos = OrderedDict(), 0
    for path, shape in spec.items():
        n = int(np.prod(shape))
        out[path] = blob[pos:pos + n].reshape(shape).copy()
        pos += n
    return out


class TrainState:
    """Parameters, gradients and Adam moments of one network as flat float32 CUDA tensors (canonical spec order),
    plus the 0-based optimiser step `step` (flax TrainState.step)."""

    def __init__(self, kind: str, spec, cfg, params: Dict[str, np.ndarray], lr_schedule: Callable[[int], float],
                 b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, device="cuda", precision="fp32"):
        self.lib = N.load()
        self.kind, self.spec, self.cfg = kind, spec, cfg
        self.precision = {"fp32": 0, "bf16": 1}[precision]
        self.lr_schedule, self.b1, self.b2, self.eps = lr_schedule, b1, b2, eps
        self.params = torch.from_numpy(P.flatten_params(spec, params)).to(device)
        self.grads = torch.zeros_like(self.params)
        self.mu = torch.zeros_like(self.params)
        self.nu = torch.zeros_like(self.params)
        self.loss = torch.zeros(1, dtype=torch.float32, device=device)
        self.step = 0
        self._h = C.c_void_p()
        create = self.lib.ldp_unet_trainer_create if kind == "planner" else self.lib.ldp_idm_trainer_create
        N.check(create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ldp_trainer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    def zero_grad(self):
        self.grads.zero_()
        self.loss.zero_()

    def planner_loss_grad(self, x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, cond: torch.Tensor,
                          weight: float = 1.0) -> torch.Tensor:
        """plan_loss and its gradient (agent/ldp_agent.py:113-126).  x0, noise (B,T,D); t (B,) int; cond (B,Dc).
        Adds weight * dloss/dparams into `grads`; returns the (unweighted) loss as a device scalar."""
        assert self.kind == "planner"
        x0, noise, cond = (v.to(torch.float32).contiguous() for v in (x0, noise, cond))
        t = t.to(torch.int32).contiguous()
        B, T, D = x0.shape
        if noise.shape != x0.shape or t.shape != (B,) or cond.shape != (B, self.cfg.global_cond_dim) or D != self.cfg.input_dim:
            raise ValueError("planner_loss_grad: shape mismatch")
        before = self.loss.clone()
        N.check(self.lib.ldp_unet_loss_grad(self._h, self.precision, self.params.data_ptr(), self.grads.data_ptr(), x0.data_ptr(),
                                            noise.data_ptr(), t.data_ptr(), cond.data_ptr(), B, T, float(weight),
                                            self.loss.data_ptr(), self._stream()))
        return (self.loss - before)[0]

    def idm_loss_grad(self, s: torch.Tensor, a0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor,
                      weight: float = 1.0) -> torch.Tensor:
        """idm_loss and its gradient (agent/ldp_agent.py:128-139).  s (N,2D); a0, noise (N,A); t (N,) or (N,1) int."""
        assert self.kind == "idm"
        s, a0, noise = (v.to(torch.float32).contiguous() for v in (s, a0, noise))
        t = t.reshape(-1).to(torch.int32).contiguous()
        n = s.shape[0]
        if s.shape != (n, 2 * self.cfg.obs_dim) or a0.shape != (n, self.cfg.action_dim) or noise.shape != a0.shape or t.shape != (n,):
            raise ValueError("idm_loss_grad: shape mismatch")
        before = self.loss.clone()
        N.check(self.lib.ldp_idm_loss_grad(self._h, self.precision, self.params.data_ptr(), self.grads.data_ptr(), s.data_ptr(),
                                           a0.data_ptr(), noise.data_ptr(), t.data_ptr(), n, float(weight),
                                           self.loss.data_ptr(), self._stream()))
        return (self.loss - before)[0]

    def apply_gradients(self, grad_scale: float = 1.0) -> float:
        """TrainState.apply_gradients with optax.adam(lr_schedule): the learning rate is the schedule at the
        pre-increment step; returns it."""
        lr = float(self.lr_schedule(self.step))
        self.step += 1
        N.check(self.lib.ldp_adam_update(self.params.data_ptr(), self.grads.data_ptr(), self.mu.data_ptr(),
                                         self.nu.data_ptr(), self.params.numel(), lr, self.b1, self.b2, self.eps,
                                         self.step, float(grad_scale), self._stream()))
        return lr

    def get_params(self) -> Dict[str, np.ndarray]:
        return unflatten_params(self.spec, self.params.detach().cpu().numpy())

    def grad_buckets(self):
        """Ranges of `grads` in the order the last loss/gradient call completes them: [(offset, length, has_event)]."""
        off = np.zeros(64, np.int64)
        ln = np.zeros(64, np.int64)
        ev = np.zeros(64, np.int32)
        n = C.c_int(0)
        N.check(self.lib.ldp_trainer_grad_buckets(self._h, off.ctypes.data, ln.ctypes.data, ev.ctypes.data, 64, C.byref(n)))
        return [(int(off[i]), int(ln[i]), bool(ev[i])) for i in range(n.value)]

    def allreduce_grads_bucketed(self, comm_stream: "torch.cuda.Stream", group=None):
        """Start the all-reduce of every gradient bucket on `comm_stream` as soon as the backward pass has finished it (the
        step records one event per bucket; buckets without an event wait for the whole call).  Returns the pending works -
        `wait()` them before the optimiser - and the 1/world factor for `apply_gradients`."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        main = torch.cuda.current_stream()
        works = []
        for k, (off, ln, has_ev) in enumerate(self.grad_buckets()):
            if ln == 0:
                continue
            if has_ev:
                N.check(self.lib.ldp_trainer_wait_bucket(self._h, k, comm_stream.cuda_stream))
            else:
                comm_stream.wait_stream(main)
            with torch.cuda.stream(comm_stream):
                works.append(dist.all_reduce(self.grads[off:off + ln], op=dist.ReduceOp.SUM, group=group, async_op=True))
        return works, 1.0 / world

    def grads_dict(self) -> Dict[str, np.ndarray]:
        return unflatten_params(self.spec, self.grads.detach().cpu().numpy())
