"""Torch-facing wrappers over the C ABI handles: device tensors in, device tensors out.

PyTorch provides device memory and the current CUDA stream; every contraction, norm and scheduler update runs
in libldp_b200's own kernels.  All functions raise on failure (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import params as P


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    return t


def _prec(precision) -> int:
    if precision in (N.PREC_FP32, "fp32", "float32", torch.float32):
        return N.PREC_FP32
    if precision in (N.PREC_BF16, "bf16", "bfloat16", torch.bfloat16):
        return N.PREC_BF16
    raise ValueError(f"unknown precision {precision!r}")


def _sampler(s) -> int:
    if s in (N.SAMPLER_DDPM, "ddpm"):
        return N.SAMPLER_DDPM
    if s in (N.SAMPLER_DDIM, "ddim"):
        return N.SAMPLER_DDIM
    raise ValueError(f"unknown sampler {s!r}")


# ------------------------------------------------------------------------------------------------
# scheduler  (mirror of diffusers FlaxDDPMScheduler as used at reference agent/ldp_agent.py:637-650)
# ------------------------------------------------------------------------------------------------
class DDPMSchedulerState:
    def __init__(self, betas, alphas, alphas_cumprod):
        self.betas, self.alphas, self.alphas_cumprod = betas, alphas, alphas_cumprod


class DDPMScheduler:
    """`FlaxDDPMScheduler(num_train_timesteps, beta_schedule='squaredcos_cap_v2', clip_sample=True,
    prediction_type='epsilon')` - same constructor keywords, same `create_state / step / add_noise` calls."""

    def __init__(self, num_train_timesteps: int = 100, beta_schedule: str = "squaredcos_cap_v2", clip_sample: bool = True,
                 prediction_type: str = "epsilon"):
        if beta_schedule != "squaredcos_cap_v2" or not clip_sample or prediction_type != "epsilon":
            raise NotImplementedError("only the configuration the reference uses is implemented "
                                      "(squaredcos_cap_v2, clip_sample=True, epsilon)")
        self.num_train_timesteps = int(num_train_timesteps)

    def create_state(self) -> DDPMSchedulerState:
        lib = N.load()
        n = self.num_train_timesteps
        b, a, c = (np.empty(n, np.float32) for _ in range(3))
        N.check(lib.ldp_ddpm_schedule(n, b.ctypes.data, a.ctypes.data, c.ctypes.data))
        return DDPMSchedulerState(b, a, c)

    def step(self, state, model_output: torch.Tensor, timestep: int, sample: torch.Tensor,
             noise: Optional[torch.Tensor] = None, seed: int = 0, stream_id: int = 0, sampler="ddpm") -> torch.Tensor:
        """`.step(state, eps, t, x, key).prev_sample`; the JAX key is replaced by either an injected N(0,1) tensor
        (`noise`) or a Philox (seed, stream_id) pair."""
        lib = N.load()
        eps, x = _f32c(model_output), _f32c(sample)
        out = torch.empty_like(x)
        z = _f32c(noise) if noise is not None else None
        N.check(lib.ldp_ddpm_step(self.num_train_timesteps, int(timestep), _sampler(sampler), eps.data_ptr(), x.data_ptr(),
                                  z.data_ptr() if z is not None else None, seed, stream_id, out.data_ptr(), x.numel(),
                                  _stream()))
        return out

    def add_noise(self, state, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        lib = N.load()
        x0, z = _f32c(original_samples), _f32c(noise)
        t = timesteps.reshape(-1).to(torch.int32).contiguous()
        rows = t.numel()
        if x0.shape[0] != rows:
            raise ValueError("timesteps must have one entry per leading row of original_samples")
        out = torch.empty_like(x0)
        N.check(lib.ldp_ddpm_add_noise(self.num_train_timesteps, x0.data_ptr(), z.data_ptr(), t.data_ptr(), out.data_ptr(),
                                       rows, x0.numel() // rows, _stream()))
        return out


def philox_normal(seed: int, stream_id: int, step: int, n: int, device="cuda") -> torch.Tensor:
    lib = N.load()
    out = torch.empty(n, dtype=torch.float32, device=device)
    N.check(lib.ldp_philox_normal(seed, stream_id, step, out.data_ptr(), n, _stream()))
    return out


def philox_normal_rows(seed: int, stream_id: int, step: int, row0: int, rows: int, row_len: int, device="cuda") -> torch.Tensor:
    """The noise the fused loops draw at reverse step `step` for global rows [row0, row0+rows)."""
    lib = N.load()
    out = torch.empty(rows, row_len, dtype=torch.float32, device=device)
    N.check(lib.ldp_philox_normal_rows(seed, stream_id, step, row0, rows, row_len, out.data_ptr(), _stream()))
    return out


def tc_dense(a: torch.Tensor, w: np.ndarray, bias: Optional[np.ndarray]) -> torch.Tensor:
    """C = A W + b on the tcgen05 path (bf16 operands, fp32 accumulate) - test / roofline helper."""
    lib = N.load()
    a = _f32c(a)
    w = np.ascontiguousarray(w, np.float32)
    M, K = a.shape
    Nn = w.shape[1]
    b = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    out = torch.empty(M, Nn, dtype=torch.float32, device=a.device)
    N.check(lib.ldp_tc_dense(a.data_ptr(), w.ctypes.data, b.ctypes.data if b is not None else None, out.data_ptr(),
                             M, K, Nn, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# planner
# ------------------------------------------------------------------------------------------------
class Planner:
    """ConditionalUnet1D + its reverse-diffusion loop (reference networks/diffusion_nets_v2.py:104-169,
    agent/ldp_agent.py:459-476)."""

    def __init__(self, params: Dict[str, np.ndarray], input_dim: int, global_cond_dim: int,
                 down_dims: Sequence[int] = (256, 512, 1024), diffusion_step_embed_dim: int = 256, kernel_size: int = 5,
                 n_groups: int = 8, n_train_steps: int = 100):
        self.lib = N.load()
        self.input_dim, self.global_cond_dim = int(input_dim), int(global_cond_dim)
        self.n_train_steps = int(n_train_steps)
        self.spec = P.unet_spec(input_dim, global_cond_dim, down_dims, kernel_size, diffusion_step_embed_dim)
        self.cfg = N.unet_config(input_dim, global_cond_dim, down_dims, diffusion_step_embed_dim, kernel_size, n_groups,
                                 n_train_steps)
        blob = P.flatten_params(self.spec, params)
        self._h = C.c_void_p()
        N.check(self.lib.ldp_planner_create(C.byref(self.cfg), blob.ctypes.data, blob.size, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ldp_planner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, sample: torch.Tensor, timestep, global_cond: torch.Tensor, precision="bf16") -> torch.Tensor:
        x, c = _f32c(sample), _f32c(global_cond)
        B, T, D = x.shape
        if D != self.input_dim or c.shape != (B, self.global_cond_dim):
            raise ValueError(f"shape mismatch: sample {tuple(x.shape)}, cond {tuple(c.shape)}")
        out = torch.empty_like(x)
        if isinstance(timestep, torch.Tensor) and timestep.numel() > 1:
            t = timestep.reshape(-1).to(device=x.device, dtype=torch.int32).contiguous()
            if t.numel() != B:
                raise ValueError("timestep tensor must have B entries")
            tp, ts = t.data_ptr(), 0
        else:
            tp, ts = None, int(timestep)
        N.check(self.lib.ldp_unet_forward(self._h, _prec(precision), x.data_ptr(), tp, ts, c.data_ptr(), B, T,
                                          out.data_ptr(), _stream()))
        return out

    def sample(self, x_T: torch.Tensor, global_cond: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
               row_offset: int = 0, n_steps: Optional[int] = None, sampler="ddpm", precision="bf16") -> torch.Tensor:
        x, c = _f32c(x_T), _f32c(global_cond)
        B, T, D = x.shape
        if D != self.input_dim or tuple(c.shape) != (B, self.global_cond_dim):      # the C side sees pointers only
            raise ValueError(f"shape mismatch: x_T {tuple(x.shape)}, cond {tuple(c.shape)}; expected (B,T,{self.input_dim}) and "
                             f"(B,{self.global_cond_dim})")
        n_steps = self.n_train_steps if n_steps is None else int(n_steps)
        z = None
        if noise is not None:
            z = _f32c(noise)
            if tuple(z.shape) != (n_steps, B, T, D):
                raise ValueError(f"noise must be (n_steps,B,T,D) = {(n_steps, B, T, D)}, got {tuple(z.shape)}")
        out = torch.empty_like(x)
        N.check(self.lib.ldp_planner_sample(self._h, _prec(precision), _sampler(sampler), x.data_ptr(), c.data_ptr(),
                                            z.data_ptr() if z is not None else None, seed, row_offset, B, T, n_steps,
                                            out.data_ptr(), _stream()))
        return out


    def sample_host(self, x_T_host: torch.Tensor, cond_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                    **kw) -> torch.Tensor:
        """Host-buffer form of `sample` (what an env-rollout caller holds, reference utils/rm_env_utils.py:168-188):
        host -> device copies, the fused reverse loop, device -> host copy of x0, all on the current stream.
        Pinned buffers make the copies asynchronous; the caller synchronises the stream before reading `out_host`."""
        x = x_T_host.to("cuda", dtype=torch.float32, non_blocking=True)
        c = cond_host.to("cuda", dtype=torch.float32, non_blocking=True)
        out = self.sample(x, c, **kw)
        if out_host is None:
            out_host = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        return out_host

    def read_activation(self, B: int, T: int, tap_id: int) -> torch.Tensor:
        """Per-layer parity hook: the bf16 activation `tap_id` of the last bf16 `forward` at (B, T) as float32 (rows, C).
        tap_id: i = ConditionalResidualBlock1D_i output, 100+l = Downsample1d_l, 200+u = Upsample1d_u, 300 = final block."""
        buf = torch.empty(B * T * 2048, dtype=torch.float32, device="cuda")
        rows, cols = C.c_int32(0), C.c_int32(0)
        N.check(self.lib.ldp_planner_read_activation(self._h, B, T, tap_id, buf.data_ptr(), buf.numel(), C.byref(rows),
                                                     C.byref(cols), _stream()))
        return buf[:rows.value * cols.value].reshape(rows.value, cols.value).clone()

    def profile_step(self, B: int, T: int, reps: int = 20):
        """Per-kernel timing of one bf16 denoising step (diagnostics): list of dicts with us, M, N, K, block_n, epilogue."""
        us = np.zeros(128, np.float32)
        meta = np.zeros(128 * 4, np.int32)
        ph = np.zeros(128 * 8, np.float32)
        n = C.c_int(0)
        N.check(self.lib.ldp_planner_profile_step(self._h, B, T, reps, us.ctypes.data, meta.ctypes.data, ph.ctypes.data, 128,
                                                  C.byref(n), _stream()))
        out = []
        for i in range(n.value):
            m, nn, kb, packed = (int(v) for v in meta[4 * i:4 * i + 4])
            out.append(dict(us=float(us[i]), M=m, N=nn, K=kb * 64, block_n=packed & 0xffff,
                            epilogue=("plain", "gn", "ddpm", "ln")[(packed >> 16) & 0xff], aux=(packed >> 24) & 1,
                            n_acc=(packed >> 25) & 7, phases=[float(v) for v in ph[8 * i:8 * i + 8]]))
        return out


# ------------------------------------------------------------------------------------------------
# inverse dynamics
# ------------------------------------------------------------------------------------------------
class Idm:
    """MLPDiffusion + its action denoising loop (reference networks/mlp_diffusion_nets.py:50-68,
    agent/ldp_agent.py:486-505)."""

    def __init__(self, params: Dict[str, np.ndarray], obs_dim: int, action_dim: int, hidden_dim: int = 256,
                 n_blocks: int = 3, time_dim: int = 256, cond_hidden: Sequence[int] = (256, 256), n_train_steps: int = 100):
        self.lib = N.load()
        self.obs_dim, self.action_dim, self.n_train_steps = int(obs_dim), int(action_dim), int(n_train_steps)
        self.spec = P.idm_spec(obs_dim, action_dim, hidden_dim, n_blocks, time_dim, cond_hidden)
        self.cfg = N.idm_config(obs_dim, action_dim, hidden_dim, n_blocks, time_dim, cond_hidden, n_train_steps)
        blob = P.flatten_params(self.spec, params)
        self._h = C.c_void_p()
        N.check(self.lib.ldp_idm_create(C.byref(self.cfg), blob.ctypes.data, blob.size, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ldp_idm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, s: torch.Tensor, a: torch.Tensor, time, precision="bf16") -> torch.Tensor:
        s, a = _f32c(s), _f32c(a)
        n = s.shape[0]
        if s.shape[1] != 2 * self.obs_dim or tuple(a.shape) != (n, self.action_dim):
            raise ValueError(f"shape mismatch: s {tuple(s.shape)}, a {tuple(a.shape)}")
        out = torch.empty_like(a)
        if isinstance(time, torch.Tensor) and time.numel() > 1:
            t = time.reshape(-1).to(device=s.device, dtype=torch.int32).contiguous()
            if t.numel() != n:
                raise ValueError("time tensor must have N entries")
            tp, ts = t.data_ptr(), 0
        else:
            tp, ts = None, int(time)
        N.check(self.lib.ldp_idm_forward(self._h, _prec(precision), s.data_ptr(), a.data_ptr(), tp, ts, n, out.data_ptr(),
                                         _stream()))
        return out

    def sample(self, s: torch.Tensor, a_T: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 0,
               row_offset: int = 0, n_steps: Optional[int] = None, sampler="ddpm", precision="bf16") -> torch.Tensor:
        s, a = _f32c(s), _f32c(a_T)
        n = s.shape[0]
        if s.dim() != 2 or s.shape[1] != 2 * self.obs_dim or tuple(a.shape) != (n, self.action_dim):   # the C side sees pointers only
            raise ValueError(f"shape mismatch: s {tuple(s.shape)}, a_T {tuple(a.shape)}; expected (N,{2 * self.obs_dim}) and "
                             f"(N,{self.action_dim})")
        n_steps = self.n_train_steps if n_steps is None else int(n_steps)
        z = None
        if noise is not None:
            z = _f32c(noise)
            if tuple(z.shape) != (n_steps, n, self.action_dim):
                raise ValueError(f"noise must be (n_steps,N,A), got {tuple(z.shape)}")
        out = torch.empty_like(a)
        N.check(self.lib.ldp_idm_sample(self._h, _prec(precision), _sampler(sampler), s.data_ptr(), a.data_ptr(),
                                        z.data_ptr() if z is not None else None, seed, row_offset, n, n_steps,
                                        out.data_ptr(), _stream()))
        return out


# ------------------------------------------------------------------------------------------------
# VAE encoder
# ------------------------------------------------------------------------------------------------
class VaeEncoder:
    """FlaxAutoencoderKL.encode(x).latent_dist.mean (reference agent/ldp_agent.py:46-64, process_sdvae_data.py:70-73)."""

    def __init__(self, params: Dict[str, np.ndarray], block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 in_channels: int = 3, latent_channels: int = 4, layers_per_block: int = 2, norm_num_groups: int = 32,
                 image_size: int = 64):
        self.lib = N.load()
        self.spec = P.vae_encoder_spec(block_out_channels, in_channels, latent_channels, layers_per_block)
        self.cfg = N.vae_config(block_out_channels, in_channels, latent_channels, layers_per_block, norm_num_groups, image_size)
        self.latent_channels = latent_channels
        self.image_size = image_size
        self.latent_hw = image_size >> (len(block_out_channels) - 1)
        blob = P.flatten_params(self.spec, params)
        self._h = C.c_void_p()
        N.check(self.lib.ldp_vae_create(C.byref(self.cfg), blob.ctypes.data, blob.size, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ldp_vae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, images: torch.Tensor, lat_min: float = 0.0, lat_max: float = 0.0, precision="bf16") -> torch.Tensor:
        """images: (B,S,S,3) uint8 0..255 or float32 in [-1,1], NHWC, on the GPU -> (B,h,w,4) float32."""
        if not images.is_cuda or images.dim() != 4:
            raise ValueError("images must be a (B,S,S,3) CUDA tensor")
        if images.dtype == torch.uint8:
            fmt, img = 0, images.contiguous()
        else:
            fmt, img = 1, _f32c(images)
        B = img.shape[0]
        out = torch.empty(B, self.latent_hw, self.latent_hw, self.latent_channels, dtype=torch.float32, device=img.device)
        N.check(self.lib.ldp_vae_encode(self._h, _prec(precision), img.data_ptr(), fmt, B, float(lat_min), float(lat_max),
                                        out.data_ptr(), _stream()))
        return out


class VaeDecoder:
    """FlaxAutoencoderKL.decode(z).sample (reference agent/ldp_agent.py:66-85 `vae_decode`; plan_viz of sample_viz :483)."""

    def __init__(self, params: Dict[str, np.ndarray], block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 out_channels: int = 3, latent_channels: int = 4, layers_per_block: int = 2, norm_num_groups: int = 32,
                 image_size: int = 64):
        self.lib = N.load()
        self.spec = P.vae_decoder_spec(block_out_channels, out_channels, latent_channels, layers_per_block)
        self.cfg = N.vae_config(block_out_channels, out_channels, latent_channels, layers_per_block, norm_num_groups, image_size)
        self.latent_channels, self.out_channels, self.image_size = latent_channels, out_channels, image_size
        self.latent_hw = image_size >> (len(block_out_channels) - 1)
        blob = P.flatten_params(self.spec, params)
        self._h = C.c_void_p()
        N.check(self.lib.ldp_vae_decoder_create(C.byref(self.cfg), blob.ctypes.data, blob.size, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ldp_vae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, latents: torch.Tensor, precision="bf16") -> torch.Tensor:
        """latents: (B,h,w,L) float32 (un-normalised), on the GPU -> (B,S,S,3) float32 NHWC."""
        if not latents.is_cuda or latents.dim() != 4 or latents.shape[1:] != (self.latent_hw, self.latent_hw, self.latent_channels):
            raise ValueError(f"latents must be a (B,{self.latent_hw},{self.latent_hw},{self.latent_channels}) CUDA tensor")
        z = _f32c(latents)
        B = z.shape[0]
        out = torch.empty(B, self.image_size, self.image_size, self.out_channels, dtype=torch.float32, device=z.device)
        N.check(self.lib.ldp_vae_decode(self._h, _prec(precision), z.data_ptr(), B, out.data_ptr(), _stream()))
        return out

