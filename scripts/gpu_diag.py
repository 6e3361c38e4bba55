"""Diagnostic sweep for a GPU trip: runs each stage of the CUDA path against the oracle, records max-abs errors
and timings without asserting, and writes gpurun_out/diag.json.  (Development aid; the gates live in tests/.)"""
import json
import os
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import ldp_oracle as O  # noqa: E402
from latent_diffusion_planning_b200 import handles as H  # noqa: E402
from latent_diffusion_planning_b200 import params as P  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
res = {}


def err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    d = (a - b).abs()
    return {"max": float(d.max()), "mean": float(d.mean()), "ref_max": float(b.abs().max())}


def stage(name):
    def deco(fn):
        t0 = time.time()
        try:
            res[name] = fn()
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        res[name + "_wall_s"] = round(time.time() - t0, 3)
        print(name, json.dumps(res[name])[:600], flush=True)
        (OUT / "diag.json").write_text(json.dumps(res, indent=1))
        return fn
    return deco


D = 265
full = {}


@stage("tc_dense")
def _():
    out = {}
    for (M, K, N) in [(128, 64, 128), (128, 128, 128), (256, 64, 256), (300, 200, 150), (1024, 1325, 256)]:
        rng = np.random.default_rng(0)
        a = rng.standard_normal((M, K)).astype(np.float32)
        w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        o = H.tc_dense(torch.tensor(a).cuda(), w, b)
        ref = torch.tensor(a).bfloat16().double() @ torch.tensor(w).bfloat16().double() + torch.tensor(b).double()
        out[f"{M}x{K}x{N}"] = err(o, ref)
    return out


@stage("unet_create")
def _():
    full["p"] = P.init_params(P.unet_spec(D, D), seed=0)
    full["planner"] = H.Planner(full["p"], D, D)
    torch.cuda.synchronize()
    return {"ok": True}


def _inputs(B, T, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, D, generator=g), torch.rand(B, D, generator=g) * 2 - 1


@stage("unet_fp32")
def _():
    x, c = _inputs(4, 8)
    taps = {}
    ref = O.unet_forward(full["p"], x, 50, c, taps=taps)
    full["ref4"] = ref
    o = full["planner"].forward(x.cuda(), 50, c.cuda(), precision="fp32")
    return err(o, ref)


@stage("unet_bf16")
def _():
    x, c = _inputs(4, 8)
    o = full["planner"].forward(x.cuda(), 50, c.cuda(), precision="bf16")
    out = {"B4": err(o, full["ref4"])}
    f = full["planner"].forward(x.cuda(), 50, c.cuda(), precision="fp32")
    out["B4_vs_fp32"] = err(o, f)
    for (B, T, k) in [(33, 8, 0), (8, 16, 99), (6, 4, 7)]:
        x, c = _inputs(B, T, seed=B)
        o = full["planner"].forward(x.cuda(), k, c.cuda(), precision="bf16")
        f = full["planner"].forward(x.cuda(), k, c.cuda(), precision="fp32")
        out[f"B{B}_T{T}_k{k}_vs_fp32"] = err(o, f)
    return out


@stage("loop_fused_vs_unfused_bf16")
def _():
    pl = full["planner"]
    B, T, n = 8, 8, 4
    x, c = _inputs(B, T, seed=31)
    z = torch.randn(n, B, T, D, generator=torch.Generator().manual_seed(7)).cuda()
    fused = pl.sample(x.cuda(), c.cuda(), noise=z, n_steps=n, precision="bf16")
    s = H.DDPMScheduler(100)
    cur = x.cuda()
    for i in range(n):
        k = n - 1 - i
        cur = s.step(None, pl.forward(cur, k, c.cuda(), precision="bf16"), k, cur, noise=z[i])
    return err(fused, cur)


@stage("loop_bf16_vs_fp32_100")
def _():
    pl = full["planner"]
    B, T, n = 16, 8, 100
    x, c = _inputs(B, T, seed=51)
    z = torch.randn(n, B, T, D, generator=torch.Generator().manual_seed(3)).cuda()
    a = pl.sample(x.cuda(), c.cuda(), noise=z, precision="bf16")
    b = pl.sample(x.cuda(), c.cuda(), noise=z, precision="fp32")
    e = err(a, b)
    e.update(mean_a=float(a.mean()), mean_b=float(b.mean()), std_a=float(a.std()), std_b=float(b.std()))
    return e


@stage("loop_timing_bf16")
def _():
    pl = full["planner"]
    out = {}
    for B in (128, 1024):
        x, c = _inputs(B, 8, seed=B)
        x, c = x.cuda(), c.cuda()
        for _ in range(2):
            pl.sample(x, c, seed=1, precision="bf16")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pl.sample(x, c, seed=1, precision="bf16")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"B{B}"] = {"ms_per_100_steps": ms, "plans_per_s": B / ms * 1e3, "tflops_useful": B * 15.93e9 / ms / 1e9}
    return out


@stage("loop_timing_bf16_nograph_step")
def _():
    pl = full["planner"]
    x, c = _inputs(1024, 8, seed=2)
    x, c = x.cuda(), c.cuda()
    for _ in range(3):
        pl.forward(x, 50, c, precision="bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pl.forward(x, 50, c, precision="bf16")
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_forward_incl_otab": e0.elapsed_time(e1) / 20}


@stage("idm")
def _():
    p = P.init_params(P.idm_spec(D, 7), seed=1)
    idm = H.Idm(p, D, 7)
    full["idm"], full["idm_p"] = idm, p
    g = torch.Generator().manual_seed(2)
    s, a = torch.rand(300, 2 * D, generator=g) * 2 - 1, torch.randn(300, 7, generator=g)
    ref = O.idm_forward(p, s, a, 50)
    out = {"fp32": err(idm.forward(s.cuda(), a.cuda(), 50, precision="fp32"), ref)}
    out["bf16"] = err(idm.forward(s.cuda(), a.cuda(), 50, precision="bf16"), ref)
    return out


@stage("idm_loop")
def _():
    idm, p = full["idm"], full["idm_p"]
    N, n = 12, 100
    g = torch.Generator().manual_seed(8)
    s, a = torch.rand(N, 2 * D, generator=g) * 2 - 1, torch.randn(N, 7, generator=g)
    z = torch.randn(n, N, 7, generator=torch.Generator().manual_seed(4))
    ref = O.idm_sample(p, O.ddpm_schedule(100), s, a, z, n)
    out = {"fp32": err(idm.sample(s.cuda(), a.cuda(), noise=z.cuda(), precision="fp32"), ref)}
    out["bf16"] = err(idm.sample(s.cuda(), a.cuda(), noise=z.cuda(), precision="bf16"), ref)
    N = 4096
    s, a = (torch.rand(N, 2 * D, generator=g) * 2 - 1).cuda(), torch.randn(N, 7, generator=g).cuda()
    for _ in range(2):
        idm.sample(s, a, seed=1, precision="bf16")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        idm.sample(s, a, seed=1, precision="bf16")
    e1.record()
    torch.cuda.synchronize()
    out["ms_per_100_steps_N4096"] = e0.elapsed_time(e1) / 3
    return out


print(json.dumps(res, indent=1)[:200])
