"""Host-side tiling decisions of the tcgen05 launches (csrc/tc_gemm.cu tc_gemm_geometry) for the shapes bench.py runs - no device needed.
out = {tiles_m, tiles_n, grid_ctas, persistent, acc_bufs, tmem_cols, n_tail, acc_stride}."""
import ctypes as C

import pytest

from latent_diffusion_planning_b200 import _native

PLAIN, GN, DDPM, LN = 0, 1, 2, 3


@pytest.fixture(scope="module")
def lib():
    return _native.load()


def geo(lib, M, N, bn, epi, pair=0, n_acc=1):
    out = (C.c_int32 * 8)()
    st = lib.ldp_tc_geometry(M, N, bn, epi, pair, n_acc, out)
    assert st == 0, lib.ldp_last_error()
    return dict(zip(("tiles_m", "tiles_n", "grid", "persistent", "acc_bufs", "tmem_cols", "n_tail", "acc_stride"), list(out)))


def test_planner_layers_are_one_wave(lib):
    # B = 1024, T = 8: level 0 (256 channels) is 64 x 2 tiles, the deep level (T = 2, 1024 channels) 16 x 8 - one tile per CTA, <= 148 CTAs
    g = geo(lib, 8192, 256, 128, GN)
    assert (g["tiles_m"], g["tiles_n"], g["grid"], g["persistent"]) == (64, 2, 128, 0)
    g = geo(lib, 2048, 1024, 128, GN, pair=1, n_acc=3)
    assert (g["tiles_m"], g["tiles_n"], g["grid"], g["persistent"]) == (16, 8, 128, 0)
    assert g["acc_stride"] == 3 * 128 and g["tmem_cols"] == 512 and g["acc_bufs"] == 1


def test_ddpm_layer_widens_its_last_tile_instead_of_a_third_wave(lib):
    # N = 265 = 2 x 128 + 9: two N tiles, the last one 16 columns wider
    g = geo(lib, 8192, 265, 128, DDPM)
    assert (g["tiles_n"], g["n_tail"], g["grid"]) == (2, 16, 128)
    assert g["acc_stride"] == 144 and g["tmem_cols"] == 256
    # remainder > 16 columns: a real third tile
    g = geo(lib, 8192, 300, 128, DDPM)
    assert (g["tiles_n"], g["n_tail"]) == (3, 0)


def test_vae_convolutions_are_persistent_pairs_with_two_accumulator_buffers(lib):
    # 592 images of 64 x 64: level 0 (128 channels) 18944 M tiles; level 3 (8 x 8, 512 channels, BN 256) 296 x 2 tiles = 2 per CTA
    g = geo(lib, 592 * 4096, 128, 128, PLAIN, pair=1)
    assert g["persistent"] == 1 and g["grid"] == 148 and g["acc_bufs"] == 2 and g["tmem_cols"] == 256 and g["tiles_m"] % 2 == 0
    g = geo(lib, 592 * 64, 512, 256, PLAIN, pair=1)
    assert (g["tiles_m"], g["tiles_n"], g["persistent"], g["acc_bufs"], g["tmem_cols"]) == (296, 2, 1, 2, 512)
    assert g["tiles_m"] * g["tiles_n"] == 4 * 148


def test_pair_rounds_odd_tile_counts_up_and_small_launches_stay_plain_grids(lib):
    g = geo(lib, 3 * 128, 128, 128, PLAIN, pair=1)
    assert g["tiles_m"] == 4 and g["persistent"] == 0 and g["grid"] == 4
    # a paired BN = 256 launch exists only in persistent form - also when the last chunk of a batch leaves it a handful of tiles;
    # the same shape as single CTAs is a plain grid
    g = geo(lib, 8 * 64, 512, 256, PLAIN, pair=1)
    assert g["persistent"] == 1 and g["grid"] == 8 and g["acc_bufs"] == 2
    g = geo(lib, 100 * 128, 256, 256, PLAIN)
    assert g["persistent"] == 0 and g["grid"] == 100 and g["acc_bufs"] == 1


def test_tmem_budget_is_enforced(lib):
    out = (C.c_int32 * 8)()
    assert lib.ldp_tc_geometry(1024, 256, 128, GN, 0, 5, out) != 0          # 5 x 128 columns > 512
    assert b"512" in lib.ldp_last_error() or "512" in str(lib.ldp_last_error())
    assert lib.ldp_tc_geometry(1024, 256, 100, GN, 0, 1, out) != 0          # block_n
