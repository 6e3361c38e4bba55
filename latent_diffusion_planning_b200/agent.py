"""LDPAgent - host-side mirror of the reference agent's sampling surface (reference agent/ldp_agent.py:28-672).

Same method names and argument meaning as the reference (`create`, `sample`, `sample_viz`, `sample_action`,
`sample_action_from_plan`, `vae_encode`, `get_obs_cond`, `get_params`, `config[...]`) plus `act` (= `sample`, the
name BASELINE.json uses).  Batches are dicts of tensors exactly like the reference's:
`{'obs': {key: (B, H, ...)}, ['actions': (B, H, A)]}` with raw pixels 0..255.  Everything numerical runs in
libldp_b200 (VAE encoder, planner loop, IDM loop); this file only does the reference's glue: normalisation constants,
concatenations, reshapes.  Training (`update`, `update_mixed`, `get_metrics`) runs through
`train.TrainState` on the same library (csrc/train.cu); under torch.distributed the gradients are all-reduced in
buckets that start as the backward pass finishes them.

`rng` replaces the JAX PRNG key: an int seed (or anything `int()` accepts).  Noise is counter-based Philox keyed by
(seed, stream, step, GLOBAL row), so a batch sharded over ranks (`row_offset`) reproduces the unsharded result.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import handles as H
from . import jax_random as JR
from . import params as P

# Philox stream ids (disjoint streams for the independent draws of one act())
STREAM_PLANNER_LOOP, STREAM_IDM_LOOP, STREAM_PLANNER_INIT, STREAM_IDM_INIT = 0, 1, 2, 3
STREAM_TRAIN_PLANNER, STREAM_TRAIN_IDM = 4, 5


# ------------------------------------------------------------------------------------------------
# normalisation (reference utils/data_utils.py:9-80)
# ------------------------------------------------------------------------------------------------
_CONST_CACHE: Dict[Any, Any] = {}


def _as_t(v, like: torch.Tensor) -> torch.Tensor:
    """Normalisation constant -> float32 tensor on `like`'s device, uploaded ONCE per (object, device): a pageable
    host-to-device copy waits for the stream to drain, and `update` used to do ten of them per step (2.9 of its 4.0 ms)."""
    scalar = np.isscalar(v)
    key = (("s", float(v)) if scalar else ("o", id(v)), str(like.device))
    hit = _CONST_CACHE.get(key)
    if hit is not None and (scalar or hit[0] is v):
        return hit[1]
    t = torch.as_tensor(np.asarray(v, dtype=np.float32), device=like.device)
    if len(_CONST_CACHE) > 4096:
        _CONST_CACHE.clear()
    _CONST_CACHE[key] = (v, t)             # keeps `v` alive, so its id cannot be recycled while the entry exists
    return t


def normalize_unnormalize(val: torch.Tensor, spec: Dict[str, Any], normalize: bool) -> torch.Tensor:
    """One entry of `normalize_unnormalize_obs` (utils/data_utils.py:24-68)."""
    if "mean" in spec:
        raise NotImplementedError("mean/std normalisation is NotImplemented in the reference too (utils/data_utils.py:30)")
    if "min" in spec:
        lo, hi = _as_t(spec["min"], val), _as_t(spec["max"], val)
        if normalize:
            return (val - lo) / (hi - lo) * 2 - 1
        out = (val + 1) / 2 * (hi - lo) + lo
        return torch.minimum(torch.maximum(out, lo), hi)
    if "clip_min" in spec:
        return torch.minimum(torch.maximum(val, _as_t(spec["clip_min"], val)), _as_t(spec["clip_max"], val))
    raise NotImplementedError(f"unknown normalisation spec {sorted(spec)}")


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range of `rank` when n independent units are split over `world` ranks (sizes differ by <= 1).
    The reference requires `batch % n_devices == 0` (train_bc.py:73); eval batches are ragged, so we do not."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LDPAgent:
    """Sampling-side LDP agent on the B200-native kernels."""

    def __init__(self, planner: H.Planner, idm: H.Idm, vae: Optional[H.VaeEncoder], obs_normalization: Dict[str, Any],
                 config: Dict[str, Any], planner_params, idm_params, precision: str = "bf16", sampler: str = "ddpm",
                 vae_decoder: Optional["H.VaeDecoder"] = None, viz: bool = False):
        self._planner, self._idm, self.vae = planner, idm, vae
        self._train: Dict[str, Any] = {}          # name -> train.TrainState, built by the first update()
        self._stale = set()                       # networks whose inference handle lags the trained parameters
        self.data_parallel = True                 # update(): all-reduce gradients when a process group is initialised
        self._side_stream = None
        self._comm_stream = None
        self._pinned: Dict[Any, list] = {}
        self._pinned_next = 0
        self.bucketed_allreduce = True            # update(): start each gradient bucket's all-reduce as the backward finishes it
        self.vae_decoder, self.viz = vae_decoder, viz
        self.obs_normalization = obs_normalization
        self.config = config
        self._planner_params, self._idm_params = planner_params, idm_params
        self.precision, self.sampler = precision, sampler
        self.use_planner = self.use_idm = True

    # inference handles; rebuilt from the trained parameters the first time they are used after an update()
    @property
    def planner(self) -> H.Planner:
        if "planner" in self._stale:
            self._planner_params = self._train["planner"].get_params()
            old = self._planner
            self._planner = H.Planner(self._planner_params, *self._net_args["planner"])
            old.close()
            self._stale.discard("planner")
        return self._planner

    @property
    def idm(self) -> H.Idm:
        if "idm" in self._stale:
            self._idm_params = self._train["idm"].get_params()
            old = self._idm
            self._idm = H.Idm(self._idm_params, *self._net_args["idm"])
            old.close()
            self._stale.discard("idm")
        return self._idm

    # ---------------------------------------------------------------- construction
    @classmethod
    def create(cls, rng, batch=None, shape_meta=None, *, name: str = "ldp", planner: Optional[dict] = None,
               idm_net: Optional[dict] = None, preprocess_time: Optional[dict] = None, cond_encoder: Optional[dict] = None,
               vae_pretrain_path: Optional[str] = None, vae_feature_dim: int = 256, use_planner: bool = True, use_idm: bool = True,
               lowdim_obs: Sequence[str] = (), rgb_obs: Sequence[str] = (), obs_normalization: Optional[dict] = None,
               data_name: str = "", obs_horizon: int = 1, pred_horizon: int = 8, action_horizon: int = 4,
               planner_n_diffusion_steps: int = 100, idm_n_diffusion_steps: int = 100, alpha_planner: float = 1.0,
               alpha_idm: float = 1.0, lr: float = 1e-4, end_lr: float = 1e-6, idm_lr: float = 1e-4, idm_end_lr: float = 1e-6,
               warmup_steps: int = 1000, decay_steps: int = 500000, update_planner_every: int = 1, update_idm_every: int = 1,
               update_idm_after: int = -1, update_planner_until: int = -1, update_planner_after: int = -1, grad_clip=None,
               # additions (not in the reference): weights, VAE topology, compute mode
               planner_params: Optional[dict] = None, idm_params: Optional[dict] = None, vae_params: Optional[dict] = None,
               vae_block_out_channels: Sequence[int] = (128, 256, 512, 512), precision: str = "bf16", sampler: str = "ddpm",
               vae_decoder_params: Optional[dict] = None, viz: bool = False, vae_image_size: int = 64,
               vae_layers_per_block: int = 2, vae_norm_num_groups: int = 32):
        """Keyword surface of the reference's `LDPAgent.create` (agent/ldp_agent.py:516-532).  `shape_meta` is the data
        config's `{'ac_dim': A, 'all_shapes': {key: [...]}}`.  Weights: pass Flax-layout trees (flat 'a/b/kernel' dicts or
        nested), otherwise they are drawn with the reference's initialisers from `rng` (no checkpoints without network)."""
        seed = int(rng)
        shape_meta = shape_meta or {}
        all_shapes = shape_meta.get("all_shapes", {})
        action_dim = int(shape_meta["ac_dim"])
        lowdim_dim = int(sum(int(np.prod(all_shapes[k])) for k in lowdim_obs))
        obs_dim = lowdim_dim + int(vae_feature_dim) * len(rgb_obs)                    # agent/ldp_agent.py:534-539
        planner = dict(planner or {})
        down_dims = tuple(planner.get("down_dims", (256, 512, 1024)))
        dsed = int(planner.get("diffusion_step_embed_dim", 256))
        ksize = int(planner.get("kernel_size", 5))
        n_groups = int(planner.get("n_groups", 8))
        cond_dim = obs_dim * obs_horizon
        uspec = P.unet_spec(obs_dim, cond_dim, down_dims, ksize, dsed)
        if planner_params is None:
            planner_params = P.init_params(uspec, seed=seed)
        else:
            planner_params = P.canonicalize_flax_names(P.unnest(planner_params) if _is_nested(planner_params) else planner_params)
        idm_net = dict(idm_net or {})
        hidden = int(idm_net.get("hidden_dim", 256))
        # reference key: idm_net.n_blocks (agent/ldp_agent.yaml:19, MLPResNet.n_blocks); `num_blocks` accepted as an alias
        n_blocks = int(idm_net.get("n_blocks", idm_net.get("num_blocks", 3)))
        # options of the reference networks that the kernels do not implement are rejected, never silently ignored
        if not idm_net.get("use_layer_norm", True):
            raise NotImplementedError("idm_net.use_layer_norm=False: the IDM kernels implement the LayerNorm variant "
                                      "(agent/ldp_agent.yaml:21) only")
        if float(idm_net.get("dropout_rate") or 0.0) > 0.0:
            raise NotImplementedError("idm_net.dropout_rate > 0 is not implemented (the reference yaml leaves it unset)")
        if not planner.get("downsample", True):
            raise NotImplementedError("planner.downsample=False is not implemented (networks/diffusion_nets_v2.py:112 default True)")
        time_dim = int((preprocess_time or {}).get("output_size", 256))
        cond_hidden = tuple((cond_encoder or {}).get("hidden_dims", (256, 256)))
        ispec = P.idm_spec(obs_dim, action_dim, hidden, n_blocks, time_dim, cond_hidden)
        if idm_params is None:
            idm_params = P.init_params(ispec, seed=seed + 1)
        else:
            idm_params = P.canonicalize_flax_names(P.unnest(idm_params) if _is_nested(idm_params) else idm_params)
        pl = H.Planner(planner_params, obs_dim, cond_dim, down_dims, dsed, ksize, n_groups, planner_n_diffusion_steps)
        idm = H.Idm(idm_params, obs_dim, action_dim, hidden, n_blocks, time_dim, cond_hidden, idm_n_diffusion_steps)
        vae = None
        if vae_pretrain_path is not None and vae_params is None and len(rgb_obs) > 0:
            # a Flax msgpack file / directory (e.g. the HF `diffusion_flax_model.msgpack`); the reference restores an orbax
            # directory here (agent/ldp_agent.py:553, model/stable_vae_model.py:119-120), which is not restated
            from . import checkpoints as CK
            vae_params, dec_from_file = CK.load_vae_flax(vae_pretrain_path, vae_block_out_channels)
            if vae_decoder_params is None:
                vae_decoder_params = dec_from_file
        if len(rgb_obs) > 0:
            vspec = P.vae_encoder_spec(vae_block_out_channels, layers_per_block=vae_layers_per_block)
            if vae_params is None:
                vae_params = P.init_params(vspec, seed=seed + 2)
            else:
                vae_params = P.unnest(vae_params) if _is_nested(vae_params) else vae_params
            vae = H.VaeEncoder(vae_params, vae_block_out_channels, layers_per_block=vae_layers_per_block,
                               norm_num_groups=vae_norm_num_groups, image_size=vae_image_size)
            lat = vae.latent_hw * vae.latent_hw * vae.latent_channels
            if lat != int(vae_feature_dim):
                raise ValueError(f"vae_feature_dim={vae_feature_dim} but the encoder produces {lat} features per frame")
        # `viz`: build the VAE decoder for plan_viz (reference vae_decode, agent/ldp_agent.py:66-85).  The reference always
        # decodes inside sample_viz; here it is a flag because decoding Ha+1 frames per plan costs more than the plan.
        vae_dec = None
        if viz and len(rgb_obs) > 0:
            dspec = P.vae_decoder_spec(vae_block_out_channels, layers_per_block=vae_layers_per_block)
            if vae_decoder_params is None:
                vae_decoder_params = P.init_params(dspec, seed=seed + 3)
            else:
                vae_decoder_params = P.unnest(vae_decoder_params) if _is_nested(vae_decoder_params) else vae_decoder_params
            vae_dec = H.VaeDecoder(vae_decoder_params, vae_block_out_channels, layers_per_block=vae_layers_per_block,
                                   norm_num_groups=vae_norm_num_groups, image_size=vae_image_size)
        config = dict(name=name, obs_horizon=obs_horizon, action_dim=action_dim, pred_horizon=pred_horizon,
                      action_horizon=action_horizon, obs_dim=obs_dim, rgb_obs=list(rgb_obs), lowdim_obs=list(lowdim_obs),
                      vae_feature_dim=int(vae_feature_dim), planner_n_diffusion_steps=planner_n_diffusion_steps,
                      idm_n_diffusion_steps=idm_n_diffusion_steps, data_name=data_name,   # agent/ldp_agent.py:653-665
                      update_planner_every=update_planner_every, update_idm_every=update_idm_every,
                      update_idm_after=update_idm_after, update_planner_until=update_planner_until,
                      update_planner_after=update_planner_after)
        agent = cls(pl, idm, vae, obs_normalization or {}, config, planner_params, idm_params, precision, sampler, vae_dec, viz)
        agent.use_planner, agent.use_idm = bool(use_planner), bool(use_idm)
        agent.alpha_planner, agent.alpha_idm = float(alpha_planner), float(alpha_idm)
        agent._net_args = dict(planner=(obs_dim, cond_dim, down_dims, dsed, ksize, n_groups, planner_n_diffusion_steps),
                               idm=(obs_dim, action_dim, hidden, n_blocks, time_dim, cond_hidden, idm_n_diffusion_steps))
        agent._opt = dict(lr=lr, end_lr=end_lr, idm_lr=idm_lr, idm_end_lr=idm_end_lr, warmup_steps=warmup_steps,
                          decay_steps=decay_steps)
        return agent

    # ---------------------------------------------------------------- reference-named pieces
    def get_params(self):
        """`{planner_params, idm_params}` as nested Flax-style trees (agent/ldp_agent.py:508-514)."""
        if "planner" in self._train:
            self._planner_params = self._train["planner"].get_params()
        if "idm" in self._train:
            self._idm_params = self._train["idm"].get_params()
        return dict(planner_params=P.nest(self._planner_params), idm_params=P.nest(self._idm_params))

    def _postprocess_obs(self, obs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """`postprocess_batch_obs` for the low-dim keys; image keys stay raw (their x/255*2-1 is fused into the encoder)."""
        norm = self.obs_normalization.get("obs", {})
        out = {}
        for k, v in obs.items():
            v = v.cuda() if not v.is_cuda else v
            if f"latent_{k}" in self.config["rgb_obs"]:
                out[k] = v
            else:
                if k not in norm:
                    raise AssertionError(f"obs_normalization keys {sorted(norm)} do not match batch key {k}")
                out[k] = normalize_unnormalize(v.to(torch.float32), norm[k], True)
        return out

    def vae_encode(self, obs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """agent/ldp_agent.py:46-64: every image key whose `latent_<key>` is an rgb_obs is replaced by its normalised
        latent features (B, H, h*w*c)."""
        norm = self.obs_normalization.get("obs", {})
        out = {}
        for k, v in obs.items():
            lk = f"latent_{k}"
            if lk not in self.config["rgb_obs"]:
                out[k] = v
                continue
            B, Hh = v.shape[:2]
            img = v.reshape(-1, *v.shape[-3:])
            if img.dtype != torch.uint8:                 # float pixels 0..255 -> [-1, 1] with the key's own constants
                spec = norm.get(k, {"min": 0, "max": 255})
                img = normalize_unnormalize(img.to(torch.float32), spec, True)
            # latent normalisation (:62): fused into the encoder when the spec is one scalar pair; a per-dimension
            # min/max (or a clip spec) is applied on the raw latents with the same code the training path uses
            lspec = norm.get(lk)
            scalar = bool(lspec) and "min" in lspec and np.ptp(np.asarray(lspec["min"])) == 0 and np.ptp(np.asarray(lspec["max"])) == 0
            lo, hi = (float(np.min(lspec["min"])), float(np.max(lspec["max"]))) if scalar else (0.0, 0.0)
            z = self.vae.encode(img.contiguous(), lat_min=lo, lat_max=hi, precision=self.precision).reshape(B, Hh, -1)
            if lspec and not scalar:
                z = normalize_unnormalize(z, lspec, True)
            out[lk] = z
        return out

    def get_obs_cond(self, obs: Dict[str, torch.Tensor]) -> torch.Tensor:
        """agent/ldp_agent.py:88-97 - including its quirk of concatenating several cameras along the horizon axis."""
        low = torch.cat([obs[k] for k in self.config["lowdim_obs"]], dim=-1).to(torch.float32)
        B, Hh = low.shape[:2]
        feats = torch.cat([obs[k] for k in self.config["rgb_obs"]], dim=1).reshape(B, Hh, -1) if self.config["rgb_obs"] else \
            low.new_zeros(B, Hh, 0)
        return torch.cat([feats, low.reshape(B, Hh, -1)], dim=-1)

    def _unnormalize_actions(self, a: torch.Tensor) -> torch.Tensor:
        spec = self.obs_normalization.get("actions")
        return normalize_unnormalize(a, spec, False) if spec else a

    # ---------------------------------------------------------------- sampling
    def _idm_loop(self, ssp: torch.Tensor, seed, row_offset: int) -> torch.Tensor:
        n, A = ssp.shape[0], self.config["action_dim"]
        steps = self.config["idm_n_diffusion_steps"]
        if JR.is_key(seed):                       # a raw JAX key: the reference's key threading and generator (:488-503)
            start_key, step_keys, _ = JR.sampling_keys(seed, steps)
            a_T = JR.normal(start_key, n * A).reshape(n, A)
            noise = JR.normal(step_keys, n * A).reshape(steps, n, A) if self.sampler == "ddpm" else None
            return self.idm.sample(ssp, a_T, noise=noise, n_steps=steps, sampler=self.sampler, precision=self.precision)
        a_T = H.philox_normal_rows(seed, STREAM_IDM_INIT, 0, row_offset, n, A)
        return self.idm.sample(ssp, a_T, seed=seed, row_offset=row_offset, n_steps=steps,
                               sampler=self.sampler, precision=self.precision)

    def vae_decode(self, feats: torch.Tensor) -> torch.Tensor:
        """agent/ldp_agent.py:66-85: the first `vae_feature_dim` features of every frame -> (h, w, 4) latents ->
        unnormalize_obs with the first rgb key's min/max -> decode -> (B, H, 3, S, S)."""
        if self.vae_decoder is None:
            raise RuntimeError("the agent was created without viz=True: no VAE decoder")
        B, Hh = feats.shape[:2]
        d = self.vae_decoder
        n = d.latent_hw * d.latent_hw * d.latent_channels
        z = feats[:, :, :n].reshape(B * Hh, d.latent_hw, d.latent_hw, d.latent_channels).to(torch.float32)
        key = self.config["rgb_obs"][0]
        norm = self.obs_normalization.get("obs", {}).get(key)
        if norm is not None:
            z = normalize_unnormalize(z.reshape(B * Hh, -1), norm, False).reshape(z.shape)
        img = d.decode(z.contiguous(), precision=self.precision)                      # (B*H, S, S, 3) NHWC
        return img.permute(0, 3, 1, 2).reshape(B, Hh, d.out_channels, d.image_size, d.image_size)

    def sample_viz(self, batch: Dict[str, Any], eval_rng, row_offset: int = 0):
        """agent/ldp_agent.py:435-506: encode -> planner loop -> plan -> IDM loop -> actions.
        Returns `(action (B, Ha, A), {'plan': (B, Ha+1, D), 'plan_viz': (B, Ha+1, 3, S, S) | None[, 'plan_mse']})`;
        `plan_viz` is decoded when the agent was created with `viz=True` (reference :483 always decodes)."""
        cfg = self.config
        obs = self.vae_encode(self._postprocess_obs(batch["obs"]))
        obs_emb = self.get_obs_cond(obs)
        B, oh, T, D = obs_emb.shape[0], cfg["obs_horizon"], cfg["pred_horizon"], cfg["obs_dim"]
        obs_cond = obs_emb[:, :oh].reshape(B, -1).contiguous()
        steps = cfg["planner_n_diffusion_steps"]
        if JR.is_key(eval_rng):
            # `eval_rng` is a raw JAX PRNG key: draw what the reference draws (jax.random threefry, its split sequence,
            # :461-476).  The draws depend on the full batch shape, so this mode is not shard-invariant.
            if row_offset:
                raise ValueError("JAX-key noise is drawn for the whole batch: row_offset must be 0")
            start_key, step_keys, seed = JR.sampling_keys(eval_rng, steps)
            x_T = JR.normal(start_key, B * T * D).reshape(B, T, D)
            noise = JR.normal(step_keys, B * T * D).reshape(steps, B, T, D) if self.sampler == "ddpm" else None
            x0 = self.planner.sample(x_T, obs_cond, noise=noise, n_steps=steps, sampler=self.sampler, precision=self.precision)
        else:
            seed = int(eval_rng)
            x_T = H.philox_normal_rows(seed, STREAM_PLANNER_INIT, 0, row_offset * T, B * T, D).reshape(B, T, D)
            x0 = self.planner.sample(x_T, obs_cond, seed=seed, row_offset=row_offset, n_steps=steps,
                                     sampler=self.sampler, precision=self.precision)
        Ha = cfg["action_horizon"]
        plan = torch.cat([obs_emb[:, oh - 1:oh], x0[:, :Ha]], dim=1)
        ssp = torch.cat([plan[:, :-1], plan[:, 1:]], dim=-1).reshape(B * Ha, 2 * D).contiguous()
        action = self._idm_loop(ssp, seed, row_offset * Ha).reshape(B, Ha, cfg["action_dim"])
        action = self._unnormalize_actions(action)
        metrics = dict(plan_viz=self.vae_decode(plan) if (self.viz and self.vae_decoder is not None) else None, plan=plan)
        if obs_emb.shape[1] > oh:                                   # training batch: ground-truth future is present
            metrics["plan_mse"] = ((x0 - obs_emb[:, oh:]) ** 2).mean()
        return action, metrics

    def sample(self, batch, eval_rng, **kw):
        return self.sample_viz(batch, eval_rng, **kw)

    act = sample

    def sample_action(self, batch, eval_rng, row_offset: int = 0) -> torch.Tensor:
        """agent/ldp_agent.py:391-430: IDM only, on consecutive ground-truth observation pairs."""
        obs = self.vae_encode(self._postprocess_obs(batch["obs"]))
        plan = self.get_obs_cond(obs)
        B, Hh, D = plan.shape
        ssp = torch.cat([plan[:, :-1], plan[:, 1:]], dim=-1).reshape(B * (Hh - 1), 2 * D).contiguous()
        a = self._idm_loop(ssp, eval_rng if JR.is_key(eval_rng) else int(eval_rng), row_offset * (Hh - 1)).reshape(B, Hh - 1, self.config["action_dim"])
        return self._unnormalize_actions(a)

    def sample_action_from_plan(self, batch, next_plan: torch.Tensor, eval_rng, row_offset: int = 0) -> torch.Tensor:
        """agent/ldp_agent.py:350-389."""
        obs = self.vae_encode(self._postprocess_obs(batch["obs"]))
        start = self.get_obs_cond(obs)
        B, Hh, D = start.shape
        ssp = torch.cat([start, next_plan.to(start)], dim=-1).reshape(B * Hh, 2 * D).contiguous()
        a = self._idm_loop(ssp, eval_rng if JR.is_key(eval_rng) else int(eval_rng), row_offset * Hh).reshape(B, Hh, self.config["action_dim"])
        return self._unnormalize_actions(a)

    def sample_sharded(self, batch, eval_rng, gather: bool = True):
        """Data-parallel `sample`: this rank computes its contiguous slice of the plans (no collective on the data path);
        with `gather` the (B, Ha, A) actions are all-gathered over NCCL so every rank returns the full batch."""
        import torch.distributed as dist
        world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
        B = next(iter(batch["obs"].values())).shape[0]
        lo, hi = shard_rows(B, rank, world)
        part = {"obs": {k: v[lo:hi] for k, v in batch["obs"].items()}}
        action, metrics = self.sample_viz(part, eval_rng, row_offset=lo)
        if not gather or world == 1:
            return action, metrics
        return gather_rows(action, B, world), metrics

    # ---------------------------------------------------------------- training (scope row N1)
    def _upload_async(self, t: torch.Tensor, device) -> torch.Tensor:
        """Host int tensor -> device through a small ring of pinned buffers (a pageable `.to(device)` stalls the host until
        the copy is staged; the training loop issues everything else asynchronously)."""
        ring = self._pinned.setdefault((tuple(t.shape), t.dtype), [])
        if len(ring) < 8:
            ring.append(torch.empty(t.shape, dtype=t.dtype, pin_memory=True))
        buf = ring[self._pinned_next % len(ring)]
        self._pinned_next += 1
        buf.copy_(t)
        return buf.to(device, non_blocking=True)

    def _train_state(self, name: str):
        """Lazily build the flat train state (TrainState.create + optax.adam(warmup_cosine), agent/ldp_agent.py:580-631)."""
        if name not in self._train:
            from . import train as TR
            o = self._opt
            if name == "planner":
                sched = TR.warmup_cosine_decay_schedule(o["end_lr"], o["lr"], o["warmup_steps"], o["decay_steps"], o["end_lr"])
                self._train[name] = TR.TrainState("planner", self._planner.spec, self._planner.cfg, self._planner_params, sched,
                                                     precision=self.precision)
            else:
                sched = TR.warmup_cosine_decay_schedule(o["idm_end_lr"], o["idm_lr"], o["warmup_steps"], o["decay_steps"],
                                                        o["idm_end_lr"])
                self._train[name] = TR.TrainState("idm", self._idm.spec, self._idm.cfg, self._idm_params, sched, precision=self.precision)
        return self._train[name]

    def _gates(self, step: int) -> Tuple[bool, bool]:
        """Which networks this step trains (agent/ldp_agent.py:229-236)."""
        c = self.config
        use_planner = bool(self.use_planner) and step % c["update_planner_every"] == 0
        use_idm = bool(self.use_idm) and step % c["update_idm_every"] == 0 and step >= c["update_idm_after"]
        upd = (c["update_planner_until"] < 0 or step < c["update_planner_until"]) and step >= c["update_planner_after"]
        return use_planner and upd, use_idm

    def _train_inputs(self, batch):
        """postprocess_batch (utils/data_utils.py:70-74) + get_obs_cond: normalised (B,H,D) embeddings and actions."""
        obs = self.vae_encode(self._postprocess_obs(batch["obs"]))
        obs_emb = self.get_obs_cond(obs)
        a = batch["actions"]
        a = (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(obs_emb.device, torch.float32)
        spec = self.obs_normalization.get("actions")
        return obs, obs_emb, (normalize_unnormalize(a, spec, True) if spec else a)

    def update(self, batch, rng, step):
        """agent/ldp_agent.py:229-277.  Returns `(agent, metrics)`; the agent is updated in place (the reference returns
        a replaced copy).  Under torch.distributed every rank passes its equal shard of the global batch
        (train_bc.py:73 `batch % n_devices == 0`); losses are means over the global batch."""
        return self._update_step(batch, batch, rng, int(step), *self._gates(int(step)))

    def update_mixed(self, batch, mixed_batch, rng, step):
        """agent/ldp_agent.py:279-327: planner on `batch`, IDM on `mixed_batch`."""
        return self._update_step(batch, mixed_batch, rng, int(step), *self._gates(int(step)))

    def get_metrics(self, batch, rng):
        """agent/ldp_agent.py:328-348: the loss metrics of `update` on `batch`, without touching the parameters."""
        _, m = self._update_step(batch, batch, rng, 0, bool(self.use_planner), bool(self.use_idm), apply=False)
        return {k: v for k, v in m.items() if k not in ("g_norm", "planner_lr", "planner_step", "idm_lr", "idm_step", "noise_diff")}

    def load_params(self, planner_params: Optional[dict] = None, idm_params: Optional[dict] = None):
        """Replace network weights (checkpoint restore, train_bc.py:204-231): Flax-layout trees, nested or flat."""
        for name, tree in (("planner", planner_params), ("idm", idm_params)):
            if tree is None:
                continue
            flat = P.canonicalize_flax_names(P.unnest(tree) if _is_nested(tree) else tree)
            handle = self._planner if name == "planner" else self._idm
            blob = torch.from_numpy(P.flatten_params(handle.spec, flat))
            if name in self._train:
                self._train[name].params.copy_(blob)
                self._stale.add(name)
            elif name == "planner":
                self._planner_params, old = flat, self._planner
                self._planner = H.Planner(flat, *self._net_args["planner"])
                old.close()
            else:
                self._idm_params, old = flat, self._idm
                self._idm = H.Idm(flat, *self._net_args["idm"])
                old.close()

    def _update_step(self, batch, idm_batch, rng, step: int, use_planner: bool, use_idm: bool, apply: bool = True):
        import torch.distributed as dist
        from . import train as TR
        cfg = self.config
        # `rng`: an integer seed (counter-based Philox draws, shard-invariant) or a raw JAX key (the reference's generator
        # and key threading, agent/ldp_agent.py:240 / :143-155 / :115 / :133)
        jax_keys = JR.update_keys(rng, use_planner, use_idm) if JR.is_key(rng) else None
        seed = 0 if jax_keys is not None else int(rng)
        oh = cfg["obs_horizon"]
        world = dist.get_world_size() if (self.data_parallel and dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank() if world > 1 else 0
        obs, obs_emb, action = self._train_inputs(batch)
        B = obs_emb.shape[0]
        metrics: Dict[str, Any] = {}
        plan_loss = idm_loss = torch.zeros((), device=obs_emb.device)
        states = []
        losses = {}

        def run_idm():
            emb_i, act_i = (obs_emb, action) if idm_batch is batch else self._train_inputs(idm_batch)[1:]
            ts = self._train_state("idm")
            ts.zero_grad()
            ssp = torch.cat([emb_i[:, oh - 1:-1], emb_i[:, oh:]], dim=-1)
            ssp = ssp.reshape(-1, ssp.shape[-1]).contiguous()
            a0 = act_i[:, :-1].reshape(-1, act_i.shape[-1]).contiguous()
            n = a0.shape[0]
            if ssp.shape[0] != n:
                raise ValueError(f"IDM pairs: {ssp.shape[0]} transitions but {n} actions (obs and actions must share their horizon)")
            if jax_keys is not None:                             # the reference's draws, for the GLOBAL batch, sliced
                t_key, z_key = jax_keys["idm"]
                t = JR.randint(t_key, n * world, 0, cfg["idm_n_diffusion_steps"])[rank * n:(rank + 1) * n]
                noise = JR.normal(z_key, n * world * a0.shape[1]).reshape(n * world, -1)[rank * n:(rank + 1) * n]
            else:
                g = torch.Generator().manual_seed(seed * 2 + 1)  # timesteps of the GLOBAL batch, then this rank's slice
                t = torch.randint(0, cfg["idm_n_diffusion_steps"], (n * world,), generator=g)[rank * n:(rank + 1) * n]
                noise = H.philox_normal_rows(seed, STREAM_TRAIN_IDM, step, rank * n, n, a0.shape[1])
            losses["idm"] = self.alpha_idm * ts.idm_loss_grad(ssp, a0, noise, t if t.is_cuda else self._upload_async(t, obs_emb.device), self.alpha_idm)
            states.append(("idm", ts))
            return ts

        def run_planner():
            ts = self._train_state("planner")
            ts.zero_grad()
            target = obs_emb[:, oh:].contiguous()
            T, D = target.shape[1], target.shape[2]
            if jax_keys is not None:
                t_key, z_key = jax_keys["planner"]
                t = JR.randint(t_key, B * world, 0, cfg["planner_n_diffusion_steps"])[rank * B:(rank + 1) * B]
                noise = JR.normal(z_key, B * world * T * D).reshape(B * world, T, D)[rank * B:(rank + 1) * B]
            else:
                g = torch.Generator().manual_seed(seed * 2 + 0)
                t = torch.randint(0, cfg["planner_n_diffusion_steps"], (B * world,), generator=g)[rank * B:(rank + 1) * B]
                noise = H.philox_normal_rows(seed, STREAM_TRAIN_PLANNER, step, rank * B * T, B * T, D).reshape(B, T, D)
            cond = obs_emb[:, :oh].reshape(B, -1).contiguous()
            losses["planner"] = self.alpha_planner * ts.planner_loss_grad(target, noise, t if t.is_cuda else self._upload_async(t, obs_emb.device), cond, self.alpha_planner)
            states.append(("planner", ts))
            return ts

        scale = 1.0
        if world == 1:
            # The two networks are independent: the (small, launch-bound) IDM step is issued first on a side stream and
            # runs under the planner's; the streams join before the optimiser.
            main = torch.cuda.current_stream()
            side = None
            if use_idm:
                if use_planner:
                    if self._side_stream is None:
                        self._side_stream = torch.cuda.Stream()
                    side = self._side_stream
                    side.wait_stream(main)
                with torch.cuda.stream(side if side is not None else main):
                    run_idm()
            if use_planner:
                run_planner()
            if side is not None:
                main.wait_stream(side)
        else:
            # Data parallel: the planner's gradient buffer (278 MB) is exchanged in ~24 MB buckets, each started on a
            # communication stream the moment the backward pass has finished it (train.cu records an event per bucket), so
            # the all-reduce runs under the remaining backward work and under the IDM step; nothing else on the wire but
            # three loss scalars.  `bucketed_allreduce = False`: one collective per network after its backward.
            pending = []
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            if use_planner:
                ts = run_planner()
                if self.bucketed_allreduce:
                    pending += ts.allreduce_grads_bucketed(self._comm_stream)[0]
                else:
                    pending.append(dist.all_reduce(ts.grads, op=dist.ReduceOp.SUM, async_op=True))
            if use_idm:
                pending.append(dist.all_reduce(run_idm().grads, op=dist.ReduceOp.SUM, async_op=True))
            for work in pending:
                work.wait()
            scale = 1.0 / world
        plan_loss, idm_loss = losses.get("planner", plan_loss), losses.get("idm", idm_loss)
        sq = torch.zeros((), device=obs_emb.device)
        for _, ts in states:
            sq = sq + (torch.linalg.vector_norm(ts.grads) * scale) ** 2
        loss = plan_loss + idm_loss
        if world > 1:
            packed = torch.stack([plan_loss, idm_loss, loss]) / world
            dist.all_reduce(packed)
            plan_loss, idm_loss, loss = packed[0], packed[1], packed[2]
        metrics.update(plan_loss=plan_loss, idm_loss=idm_loss, loss=loss, emb_min=obs_emb.min(), emb_max=obs_emb.max(),
                       emb_mean=obs_emb.mean(), emb_std=obs_emb.std(unbiased=False), action_min=action.min(),
                       action_max=action.max())
        for k, v in obs.items():
            metrics[f"{k}_min"], metrics[f"{k}_max"] = v.min(), v.max()
        metrics["g_norm"] = sq.sqrt()
        # the reference reports both learning rates through the LAST schedule it built, the IDM's (agent/ldp_agent.py:258,619)
        report = self._train_state("idm").lr_schedule if self.use_idm else self._train_state("planner").lr_schedule
        for name in ("planner", "idm"):
            ts = dict(states).get(name)
            if ts is None:
                metrics[f"{name}_lr"], metrics[f"{name}_step"] = 0, 0
                if name == "planner":
                    metrics["noise_diff"] = 0
                continue
            metrics[f"{name}_lr"], metrics[f"{name}_step"] = report(ts.step), ts.step
            if apply:
                ts.apply_gradients(grad_scale=scale)
                self._stale.add(name)
        return self, metrics


def gather_rows(local: torch.Tensor, n_total: int, world: int) -> torch.Tensor:
    """All-gather row shards of unequal size (`shard_rows` partition) into the full (n_total, ...) tensor."""
    import torch.distributed as dist
    sizes = [shard_rows(n_total, r, world)[1] - shard_rows(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


def _is_nested(tree: dict) -> bool:
    return any(isinstance(v, dict) for v in tree.values())
