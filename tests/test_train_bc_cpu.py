"""Host pieces of the training entry point mirror (train_bc.py / utils/py_utils.py / data/robomimic_latent_data.py)."""
import csv

import numpy as np
import pytest
import torch

from latent_diffusion_planning_b200 import train_bc as TB


def test_every_matches_reference_semantics():
    e = TB.Every(10)
    assert [s for s in range(1, 31) if e(s)] == [10, 20, 30] and e(0)
    assert not any(TB.Every(-1)(s) for s in range(5)) and not TB.Every(None)(3)
    assert [s for s in range(1, 11) if TB.Every(10, action_repeat=2)(s)] == [5, 10]


def test_timer_tick_tock():
    t = TB.Timer()
    t.tick("a"); t.tock("a"); t.tick("a"); t.tock("a")
    with pytest.raises(ValueError):
        t.tock("a")
    t.tick("b")
    with pytest.raises(ValueError):
        t.tick("b")
    avg = t.get_average_times()
    assert set(avg) == {"a"} and avg["a"] >= 0 and t.counts == {}


def _episodes():
    eps = {}
    for d, n in enumerate([5, 3]):
        base = 100 * (d + 1)
        eps[f"demo_{d}"] = {"obs": {"z": (base + np.arange(n, dtype=np.float32))[:, None] * np.ones((1, 2), np.float32),
                                    "q": (base + np.arange(n, dtype=np.float32))[:, None]},
                            "actions": (base + np.arange(n, dtype=np.float32))[:, None] * np.ones((1, 3), np.float32)}
    return eps


def test_window_sampler_edge_padding():
    """data/robomimic_latent_data.py:116-147: clip the window to the demo, repeat the first / last frame."""
    ds = TB.LatentSequenceDataset(_episodes(), ["z", "q"], seq_length=4, n_frame_stack=2)
    assert len(ds) == 8
    it = ds.get_item(0)                                   # first step of demo 0: one frame of start padding
    assert it["obs"]["q"][:, 0].tolist() == [100, 100, 101, 102, 103] and it["actions"][:, 0].tolist() == [100, 101, 102, 103]
    it = ds.get_item(3)                                   # window runs past the end of demo 0 (length 5)
    assert it["obs"]["q"][:, 0].tolist() == [102, 103, 104, 104, 104] and it["actions"][:, 0].tolist() == [103, 104, 104, 104]
    it = ds.get_item(5)                                   # first step of demo 1: never reaches back into demo 0
    assert it["obs"]["z"][:, 1].tolist() == [200, 200, 201, 202, 202] and it["obs"]["z"].shape == (5, 2)
    it = ds.get_item(7)                                   # last step of demo 1
    assert it["actions"][:, 2].tolist() == [202, 202, 202, 202]
    with pytest.raises(ValueError):
        bad = _episodes()
        bad["demo_0"]["obs"]["q"] = bad["demo_0"]["obs"]["q"][:-1]
        TB.LatentSequenceDataset(bad, ["z", "q"], 4)


def test_sample_batch_shards_one_global_draw():
    ds = TB.LatentSequenceDataset(_episodes(), ["z", "q"], seq_length=3)
    full = ds.sample_batch(8, np.random.default_rng(5))
    parts = [ds.sample_batch(8, np.random.default_rng(5), rank=r, world=2) for r in range(2)]
    assert full["actions"].shape == (8, 3, 3) and parts[0]["obs"]["z"].shape == (4, 3, 2)
    assert np.array_equal(np.concatenate([p["actions"].numpy() for p in parts]), full["actions"].numpy())
    with pytest.raises(AssertionError):
        ds.sample_batch(7, np.random.default_rng(0), 0, 2)


def test_csv_logger_averages_between_dumps(tmp_path):
    lg = TB.CSVLogger(tmp_path)
    lg.log_metrics({"loss": 2.0, "lr": 1e-3}, 10)
    lg.log_metrics({"loss": 4.0, "lr": 1e-3, "skip": None}, 20)
    row = lg.dump(20)
    assert row["loss"] == 3.0 and row["step"] == 20
    lg.log_metrics({"loss": 1.0, "extra": 5.0}, 30)
    lg.dump(30)
    rows = list(csv.DictReader(open(tmp_path / "train.csv")))
    assert [r["step"] for r in rows] == ["20", "30"] and rows[1]["extra"] == "5.0" and rows[0]["extra"] == ""


@pytest.mark.parametrize("fs,seq", [(1, 4), (2, 4), (3, 2), (1, 9)])
def test_vectorised_gather_equals_per_item_windows(fs, seq):
    """The one-gather-per-key batch (clamped row indices) is exactly the reference's clip + edge padding, for every
    sample of every demo, including windows longer than the demo."""
    ds = TB.LatentSequenceDataset(_episodes(), ["z", "q"], seq_length=seq, n_frame_stack=fs)
    got = ds.gather_batch(np.arange(len(ds)))
    for i in range(len(ds)):
        it = ds.get_item(i)
        assert np.array_equal(got["actions"][i].numpy(), it["actions"])
        for k in ("z", "q"):
            assert np.array_equal(got["obs"][k][i].numpy(), it["obs"][k])
    a = ds.sample_batch(6, np.random.default_rng(3), rank=1, world=2)
    b = ds.to("cpu").sample_batch_fast(6, np.random.default_rng(3), rank=1, world=2)
    assert np.array_equal(a["actions"].numpy(), b["actions"].numpy()) and np.array_equal(a["obs"]["z"].numpy(), b["obs"]["z"].numpy())


def test_dataset_from_latent_file_roundtrip(tmp_path):
    """process_sdvae_data writer -> training dataset reader: latents flattened per frame, low-dim keys and actions
    extended to T_ep + 1 rows the way the reference does (data/robomimic_latent_data.py:95-110)."""
    from latent_diffusion_planning_b200 import process_sdvae_data as PS
    g = np.random.default_rng(0)
    eps, tables = {}, {}
    for d, n in enumerate([4, 6]):
        eps[f"demo_{d}"] = {"obs": {"q": g.normal(size=(n, 3)).astype(np.float32)}, "next_obs": {"q": g.normal(size=(n, 3)).astype(np.float32)},
                            "actions": g.normal(size=(n, 2)).astype(np.float32)}
        tables[f"data/demo_{d}/latent/cam"] = g.normal(size=(n + 1, 2, 2, 4)).astype(np.float32)
    tables.update({"data.attrs/total": np.asarray(2), "data.attrs/min_z": np.asarray(-1.0), "data.attrs/max_z": np.asarray(1.0)})
    path = PS.write_latents(tables, tmp_path, prefer_hdf5=False)
    ds = TB.LatentSequenceDataset.from_latent_file(path, eps, ["cam"], ["q"], seq_length=3)
    assert len(ds) == 5 + 7 and ds.obs_keys == ["latent_cam", "q"]
    it = ds.get_item(4)                                     # last row of demo_0: the appended next_obs / repeated action
    assert np.array_equal(it["obs"]["q"][0], eps["demo_0"]["next_obs"]["q"][-1])
    assert np.array_equal(it["actions"][0], eps["demo_0"]["actions"][-1])
    assert np.array_equal(it["obs"]["latent_cam"][0], tables["data/demo_0/latent/cam"][4].reshape(-1))
    assert it["obs"]["latent_cam"].shape == (3, 16)


class _FakeAgent:
    """Duck-typed stand-in for LDPAgent: the Workspace loop's host logic (cadence, logging, snapshots, eval averaging)
    is exercised on CPU; the kernels behind the real agent are covered by the -m gpu tests."""

    def __init__(self):
        self.config = {"obs_horizon": 1, "name": "fake"}
        self.use_planner = False
        self.calls = []
        self.w = {"planner_params": {"Dense_0": {"kernel": np.zeros((2, 2), np.float32)}},
                  "idm_params": {"Dense_0": {"bias": np.zeros(3, np.float32)}}}

    def update(self, batch, rng, step):
        self.calls.append((step, int(rng), tuple(batch["actions"].shape)))
        self.w["planner_params"]["Dense_0"]["kernel"] = self.w["planner_params"]["Dense_0"]["kernel"] + 1
        return self, {"loss": torch.tensor(10.0 - step), "planner_lr": 1e-4, "planner_step": step}

    def get_metrics(self, batch, rng):
        return {"loss": torch.tensor(2.0)}

    def sample_action(self, batch, rng):
        return batch["actions"][:, :-1] + 0.5

    def get_params(self):
        return self.w

    def load_params(self, planner_params=None, idm_params=None):
        self.loaded = (planner_params, idm_params)


def test_workspace_loop_cadence_snapshots_and_eval_on_cpu(tmp_path):
    ds = TB.LatentSequenceDataset(_episodes(), ["z", "q"], seq_length=3)
    ag = _FakeAgent()
    ws = TB.Workspace(ag, ds, tmp_path, eval_dataset=ds, seed=1, batch_size=4, n_grad_steps=7, log_every_step=2, dump_every_step=4,
                      save_every_step=3, eval_every_step=6, n_eval_batches=1)
    last = ws.run()
    assert [c[0] for c in ag.calls] == list(range(7)) and all(c[2] == (4, 3, 3) for c in ag.calls)
    assert len({c[1] for c in ag.calls}) == 7                                  # a fresh update seed every step
    rows = list(csv.DictReader(open(tmp_path / "train.csv")))
    assert [r["step"] for r in rows] == ["4"]                                  # dumps at multiples of 4 within 7 steps
    assert float(rows[0]["loss"]) == pytest.approx(np.mean([10.0 - 1, 10.0 - 3]))   # logged at steps 2 and 4 (metrics of update 1, 3)
    assert sorted(p.name for p in (tmp_path / "ckpt").iterdir()) == ["3.ckpt.npz", "6.ckpt.npz"]
    ev = list(csv.DictReader(open(tmp_path / "eval.csv")))
    assert len(ev) == 1 and float(ev[0]["evaldata/action_mse"]) == pytest.approx(0.25) and float(ev[0]["evaldata/action_l1"]) == pytest.approx(0.5)
    ws2 = TB.Workspace(_FakeAgent(), ds, tmp_path / "b")
    ws2.load_snapshot(tmp_path / "ckpt" / "6.ckpt.npz")
    assert ws2.step == 6 and np.array_equal(ws2.agent.loaded[0]["Dense_0/kernel"], np.full((2, 2), 6.0, np.float32))
    assert set(ws2.agent.loaded[1]) == {"Dense_0/bias"}
    assert float(last["loss"]) == 4.0 or torch.is_tensor(last["loss"])
