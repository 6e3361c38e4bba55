"""ctypes binding of libldp_b200.so (C ABI declared in include/ldp_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
PyTorch is used by callers only for device memory and streams; nothing here takes a torch type - pointers
and sizes are passed as integers.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libldp_b200.so"

LDP_OK = 0
PREC_FP32, PREC_BF16 = 0, 1
SAMPLER_DDPM, SAMPLER_DDIM = 0, 1

EXPORTS = [
    "ldp_last_error", "ldp_version", "ldp_device_check",
    "ldp_ddpm_schedule", "ldp_ddpm_step", "ldp_ddpm_add_noise", "ldp_philox_normal", "ldp_philox_normal_rows",
    "ldp_planner_create", "ldp_planner_destroy", "ldp_unet_param_count", "ldp_unet_forward", "ldp_planner_sample",
    "ldp_planner_profile_step", "ldp_planner_read_activation",
    "ldp_idm_create", "ldp_idm_destroy", "ldp_idm_param_count", "ldp_idm_forward", "ldp_idm_sample",
    "ldp_vae_create", "ldp_vae_destroy", "ldp_vae_param_count", "ldp_vae_encode",
    "ldp_vae_decoder_create", "ldp_vae_decoder_param_count", "ldp_vae_decode",
    "ldp_unet_trainer_create", "ldp_idm_trainer_create", "ldp_trainer_destroy", "ldp_unet_loss_grad",
    "ldp_idm_loss_grad", "ldp_adam_update", "ldp_trainer_grad_buckets", "ldp_trainer_wait_bucket",
    "ldp_jax_random",
    "ldp_tc_dense", "ldp_tc_geometry", "ldp_launch_count", "ldp_launch_count_reset",
]


class LdpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libldp_b200 error {code}: {msg}")
        self.code = code


class UnetConfig(C.Structure):
    _fields_ = [("input_dim", C.c_int32), ("global_cond_dim", C.c_int32), ("step_embed_dim", C.c_int32),
                ("n_levels", C.c_int32), ("down_dims", C.c_int32 * 6), ("kernel_size", C.c_int32),
                ("n_groups", C.c_int32), ("n_train_steps", C.c_int32)]


class IdmConfig(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("action_dim", C.c_int32), ("hidden_dim", C.c_int32), ("n_blocks", C.c_int32),
                ("time_dim", C.c_int32), ("n_cond_layers", C.c_int32), ("cond_hidden", C.c_int32 * 4),
                ("n_train_steps", C.c_int32)]


class VaeConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("latent_channels", C.c_int32), ("n_blocks", C.c_int32),
                ("block_out_channels", C.c_int32 * 8), ("layers_per_block", C.c_int32),
                ("norm_num_groups", C.c_int32), ("image_size", C.c_int32)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built - there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -m latent_diffusion_planning_b200.build` (needs nvcc; cross-compiles for sm_100a without a GPU).")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64, u32, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float
    lib.ldp_last_error.restype = C.c_char_p
    lib.ldp_last_error.argtypes = []
    lib.ldp_version.restype = i32
    lib.ldp_device_check.restype = i32
    lib.ldp_ddpm_schedule.argtypes = [i32, vp, vp, vp]
    lib.ldp_ddpm_step.argtypes = [i32, i32, i32, vp, vp, vp, u64, u32, vp, i64, vp]
    lib.ldp_ddpm_add_noise.argtypes = [i32, vp, vp, vp, vp, i64, i64, vp]
    lib.ldp_philox_normal.argtypes = [u64, u32, u32, vp, i64, vp]
    lib.ldp_philox_normal_rows.argtypes = [u64, u32, u32, i64, i64, i32, vp, vp]
    lib.ldp_planner_create.argtypes = [C.POINTER(UnetConfig), vp, u64, C.POINTER(vp)]
    lib.ldp_planner_destroy.argtypes = [vp]
    lib.ldp_unet_param_count.argtypes = [C.POINTER(UnetConfig)]
    lib.ldp_unet_param_count.restype = i64
    lib.ldp_unet_forward.argtypes = [vp, i32, vp, vp, i32, vp, i32, i32, vp, vp]
    lib.ldp_planner_sample.argtypes = [vp, i32, i32, vp, vp, vp, u64, i64, i32, i32, i32, vp, vp]
    lib.ldp_planner_profile_step.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.ldp_planner_read_activation.argtypes = [vp, i32, i32, i32, vp, i64, vp, vp, vp]
    lib.ldp_idm_create.argtypes = [C.POINTER(IdmConfig), vp, u64, C.POINTER(vp)]
    lib.ldp_idm_destroy.argtypes = [vp]
    lib.ldp_idm_param_count.argtypes = [C.POINTER(IdmConfig)]
    lib.ldp_idm_param_count.restype = i64
    lib.ldp_idm_forward.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp, vp]
    lib.ldp_idm_sample.argtypes = [vp, i32, i32, vp, vp, vp, u64, i64, i32, i32, vp, vp]
    lib.ldp_vae_create.argtypes = [C.POINTER(VaeConfig), vp, u64, C.POINTER(vp)]
    lib.ldp_vae_destroy.argtypes = [vp]
    lib.ldp_vae_param_count.argtypes = [C.POINTER(VaeConfig)]
    lib.ldp_vae_param_count.restype = i64
    lib.ldp_vae_encode.argtypes = [vp, i32, vp, i32, i32, f32, f32, vp, vp]
    lib.ldp_vae_decoder_create.argtypes = [C.POINTER(VaeConfig), vp, u64, C.POINTER(vp)]
    lib.ldp_vae_decoder_param_count.argtypes = [C.POINTER(VaeConfig)]
    lib.ldp_vae_decoder_param_count.restype = i64
    lib.ldp_vae_decode.argtypes = [vp, i32, vp, i32, vp, vp]
    lib.ldp_unet_trainer_create.argtypes = [C.POINTER(UnetConfig), C.POINTER(vp)]
    lib.ldp_idm_trainer_create.argtypes = [C.POINTER(IdmConfig), C.POINTER(vp)]
    lib.ldp_trainer_destroy.argtypes = [vp]
    lib.ldp_unet_loss_grad.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, vp]
    lib.ldp_idm_loss_grad.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp]
    lib.ldp_trainer_grad_buckets.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.ldp_trainer_wait_bucket.argtypes = [vp, i32, vp]
    lib.ldp_adam_update.argtypes = [vp, vp, vp, vp, u64, f32, f32, f32, f32, i64, f32, vp]
    lib.ldp_jax_random.argtypes = [vp, i32, i64, i32, vp, vp]
    lib.ldp_tc_dense.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.ldp_tc_geometry.argtypes = [i32, i32, i32, i32, i32, i32, vp]
    lib.ldp_launch_count.restype = i64
    lib.ldp_launch_count_reset.restype = None
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ldp_version",):
            pass
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != LDP_OK:
        raise LdpError(status, (load().ldp_last_error() or b"").decode("utf-8", "replace"))


def unet_config(input_dim: int, global_cond_dim: int, down_dims: Sequence[int] = (256, 512, 1024),
                step_embed_dim: int = 256, kernel_size: int = 5, n_groups: int = 8, n_train_steps: int = 100) -> UnetConfig:
    cfg = UnetConfig()
    cfg.input_dim, cfg.global_cond_dim, cfg.step_embed_dim = input_dim, global_cond_dim, step_embed_dim
    cfg.n_levels = len(down_dims)
    for i, d in enumerate(down_dims):
        cfg.down_dims[i] = d
    cfg.kernel_size, cfg.n_groups, cfg.n_train_steps = kernel_size, n_groups, n_train_steps
    return cfg


def idm_config(obs_dim: int, action_dim: int, hidden_dim: int = 256, n_blocks: int = 3, time_dim: int = 256,
               cond_hidden: Sequence[int] = (256, 256), n_train_steps: int = 100) -> IdmConfig:
    cfg = IdmConfig()
    cfg.obs_dim, cfg.action_dim, cfg.hidden_dim, cfg.n_blocks, cfg.time_dim = obs_dim, action_dim, hidden_dim, n_blocks, time_dim
    cfg.n_cond_layers = len(cond_hidden)
    for i, d in enumerate(cond_hidden):
        cfg.cond_hidden[i] = d
    cfg.n_train_steps = n_train_steps
    return cfg


def vae_config(block_out_channels: Sequence[int] = (128, 256, 512, 512), in_channels: int = 3, latent_channels: int = 4,
               layers_per_block: int = 2, norm_num_groups: int = 32, image_size: int = 64) -> VaeConfig:
    cfg = VaeConfig()
    cfg.in_channels, cfg.latent_channels, cfg.n_blocks = in_channels, latent_channels, len(block_out_channels)
    for i, d in enumerate(block_out_channels):
        cfg.block_out_channels[i] = d
    cfg.layers_per_block, cfg.norm_num_groups, cfg.image_size = layers_per_block, norm_num_groups, image_size
    return cfg
