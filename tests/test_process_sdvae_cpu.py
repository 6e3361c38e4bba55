"""Host logic of the process_sdvae_data mirror (reference process_sdvae_data.py:52-118): episode assembly, sharding,
min/max bookkeeping and the on-disk keys, with a stand-in encoder (no GPU)."""
import numpy as np
import pytest

from latent_diffusion_planning_b200 import process_sdvae_data as PS


def _episodes(lengths, seed=0):
    g = np.random.default_rng(seed)
    eps = {}
    for i, n in enumerate(lengths):
        obs = g.integers(0, 256, size=(n, 8, 8, 3), dtype=np.uint8)
        nxt = g.integers(0, 256, size=(n, 8, 8, 3), dtype=np.uint8)
        eps[f"demo_{i}"] = {"obs": {"agentview_image": obs}, "next_obs": {"agentview_image": nxt}}
    return eps


def _fake_encode(calls):
    def enc(frames):
        calls.append(len(frames))
        x = frames.astype(np.float32) / 255 * 2 - 1
        return x.reshape(len(frames), 2, 4, 2, 4, 3).mean(axis=(2, 4))[..., [0, 1, 2, 0]]      # (n, 2, 2, 4)
    return enc


def test_rm_appends_last_next_obs_and_shards():
    eps = _episodes([5, 3])
    calls = []
    t = PS.encode_dataset(eps, ["agentview_image"], _fake_encode(calls), "rm_lift", shard=4)
    assert t["data/demo_0/latent/agentview_image"].shape == (6, 2, 2, 4)         # T_ep + 1 frames
    assert t["data/demo_1/latent/agentview_image"].shape == (4, 2, 2, 4)
    assert calls == [4, 2, 4]                                                    # ragged last shard, no padding
    last = PS.episode_frames(eps["demo_0"], "agentview_image", "rm_lift")[-1]
    assert np.array_equal(last, eps["demo_0"]["next_obs"]["agentview_image"][-1])
    assert t["data.attrs/total"] == 2


def test_aloha_uses_obs_only_and_minmax_start_at_zero():
    eps = _episodes([4])
    t = PS.encode_dataset(eps, ["agentview_image"], lambda f: np.full((len(f), 1, 1, 4), 0.25, np.float32), "aloha_cube", shard=16)
    assert t["data/demo_0/latent/agentview_image"].shape[0] == 4
    assert float(t["data.attrs/min_z"]) == 0.0 and float(t["data.attrs/max_z"]) == 0.25      # min starts at 0 (reference line 71)


def test_empty_dataset_and_bad_names(tmp_path):
    t = PS.encode_dataset({}, ["agentview_image"], lambda f: f, "rm_lift")
    assert int(t["data.attrs/total"]) == 0 and float(t["data.attrs/min_z"]) == 0.0
    with pytest.raises(ValueError):
        PS.encode_dataset(_episodes([2]), ["agentview_image"], lambda f: f, "kitchen")
    with pytest.raises(ValueError):
        PS.encode_dataset(_episodes([2]), ["agentview_image"], lambda f: f, "rm_lift", shard=0)


def test_roundtrip_file(tmp_path):
    eps = _episodes([3, 2], seed=1)
    t = PS.encode_dataset(eps, ["agentview_image"], _fake_encode([]), "rm_lift", shard=8)
    path = PS.write_latents(t, tmp_path, prefer_hdf5=False)
    assert path.name == "latent.npz"
    st = PS.read_latent_stats(path)
    assert st["total"] == 2 and st["min_z"] <= 0.0 <= st["max_z"]
    with np.load(path) as z:
        assert np.array_equal(z["data/demo_1/latent/agentview_image"], t["data/demo_1/latent/agentview_image"])
