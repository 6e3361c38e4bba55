"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/ldp_b200.h declares; host-only entry points (schedule) match the oracle bit for bit; argument errors are
reported through status codes, not crashes.  No GPU compute is attempted here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np

from oracle import ldp_oracle as O
from latent_diffusion_planning_b200 import _native as N
from latent_diffusion_planning_b200 import params as P

HEADER = Path(__file__).resolve().parents[1] / "include" / "ldp_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return re.findall(r"LDP_API\s+[\w\s\*]+?\b(ldp_\w+)\s*\(", text)


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ldp_b200.h but not exported by libldp_b200.so"
    assert set(N.EXPORTS) == set(names)


def test_schedule_host_matches_oracle(lib):
    for n in (100, 10, 1000):
        b, a, c = (np.empty(n, np.float32) for _ in range(3))
        assert lib.ldp_ddpm_schedule(n, b.ctypes.data, a.ctypes.data, c.ctypes.data) == 0
        ob, oa, oc = O.ddpm_schedule(n)
        assert np.array_equal(b, ob) and np.array_equal(a, oa) and np.array_equal(c, oc)


def test_param_counts_match_specs(lib):
    cfg = N.unet_config(265, 265)
    assert lib.ldp_unet_param_count(C.byref(cfg)) == P.spec_size(P.unet_spec(265, 265)) == 69480457
    cfg = N.unet_config(25, 50, (64, 128))
    assert lib.ldp_unet_param_count(C.byref(cfg)) == P.spec_size(P.unet_spec(25, 50, (64, 128)))
    icfg = N.idm_config(265, 7)
    assert lib.ldp_idm_param_count(C.byref(icfg)) == P.spec_size(P.idm_spec(265, 7)) == 1914887
    icfg = N.idm_config(30, 14)
    assert lib.ldp_idm_param_count(C.byref(icfg)) == P.spec_size(P.idm_spec(30, 14))


def test_argument_errors_are_status_codes(lib):
    assert lib.ldp_ddpm_schedule(0, None, None, None) == -1
    assert b"bad arguments" in lib.ldp_last_error()
    bad = N.unet_config(265, 265, kernel_size=3)
    assert lib.ldp_unet_param_count(C.byref(bad)) == -1
    h = C.c_void_p()
    assert lib.ldp_planner_create(C.byref(N.unet_config(265, 265)), None, 0, C.byref(h)) == -1
    assert lib.ldp_version() >= 100


def test_flatten_roundtrip():
    spec = P.idm_spec(25, 7)
    p = P.init_params(spec, 3)
    blob = P.flatten_params(spec, p)
    assert blob.dtype == np.float32 and blob.size == P.spec_size(spec)
    off = 0
    for k, shp in spec.items():
        n = int(np.prod(shp))
        assert np.array_equal(blob[off:off + n].reshape(shp), p[k])
        off += n
    assert P.unnest(P.nest(p)).keys() == p.keys()


def test_vae_param_counts_match_specs(lib):
    """Encoder and decoder (next-row N2) blobs: the C side's parameter walk and params.py's spec agree on the size;
    SD-VAE [128,256,512,512]: 34.16 M encoder / 49.49 M decoder parameters."""
    for blocks, lpb in (((128, 256, 512, 512), 2), ((128, 256, 256, 256, 256, 256), 2), ((32, 64), 1)):
        cfg = N.vae_config(blocks, 3, 4, lpb, 32 if blocks[0] >= 128 else 8, 64 if blocks[0] >= 128 else 16)
        assert lib.ldp_vae_param_count(C.byref(cfg)) == P.spec_size(P.vae_encoder_spec(blocks, 3, 4, lpb))
        assert lib.ldp_vae_decoder_param_count(C.byref(cfg)) == P.spec_size(P.vae_decoder_spec(blocks, 3, 4, lpb))
    assert round(P.spec_size(P.vae_decoder_spec()) / 1e6, 2) == 49.49


def test_oracle_decoder_upsample_and_shapes():
    """Nearest x2 upsampling index map (out[i] = in[i // 2]) and the decoder's output geometry on integer-valued inputs."""
    import torch
    x = torch.arange(2 * 3 * 3 * 1, dtype=torch.float64).reshape(2, 3, 3, 1)
    up = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    for i in range(6):
        for j in range(6):
            assert float(up[1, i, j, 0]) == float(x[1, i // 2, j // 2, 0])
    blocks = (32, 64, 64)
    p = P.init_params(P.vae_decoder_spec(blocks, 3, 4, 1), seed=0)
    y = O.vae_decode(p, torch.zeros(1, 4, 4, 4), blocks, 1, 8)
    assert tuple(y.shape) == (1, 16, 16, 3)


def test_training_and_rng_entry_points_reject_bad_arguments(lib):
    """Row N1 / N4 exports: argument checks fire before any CUDA call, so they are testable without a GPU."""
    assert lib.ldp_adam_update(None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == -1
    assert b"bad arguments" in lib.ldp_last_error()
    assert lib.ldp_unet_loss_grad(None, 0, None, None, None, None, None, None, 1, 8, 1.0, None, None) == -1
    assert b"trainer handle" in lib.ldp_last_error()
    assert lib.ldp_idm_loss_grad(None, 0, None, None, None, None, None, None, 1, 1.0, None, None) == -1
    assert lib.ldp_unet_trainer_create(None, None) == -1
    assert lib.ldp_jax_random(None, 1, 4, 1, None, None) == -1
    assert lib.ldp_jax_random(None, 1, 4, 7, None, None) == -1 and b"mode" in lib.ldp_last_error()
    assert lib.ldp_trainer_destroy(None) == 0
