#!/usr/bin/env python
"""bench.py - the planner reverse-diffusion loop (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: 100 DDPM reverse steps of the planner score network
(ConditionalUnet1D, D = 8*8*4 + 9 = 265) on B = 1024 plans of horizon 9 (T = 8), bf16 tensor-core path, the scheduler
update fused into the last GEMM, in-kernel Philox noise.  Weak scaling: every rank runs its own 1024 plans
(independent units, no data-path collective); `value` = plans of all ranks / max-over-ranks device time.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how each field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

B_PLANS, T_PRED, LATENT, LOWDIM, N_DIFF = 1024, 8, 8 * 8 * 4, 9, 100
D_OBS = LATENT + LOWDIM
METRIC, UNIT = "planner_plans_per_sec", "plans/s"


# ----------------------------------------------------------------------------------------------------
# algorithmic work: useful FLOPs = 2*MAC without the multiplies against zero padding and without
# the step-/batch-invariant FiLM work (hoisted out of the loop)
# ----------------------------------------------------------------------------------------------------
def unet_useful_flops(D: int, T: int, down=(256, 512, 1024), k: int = 5) -> int:
    def k5(t):
        return sum(1 for tt in range(t) for j in range(k) if 0 <= tt + j - k // 2 < t)

    def dn(t):
        return sum(1 for tt in range(t // 2) for j in range(3) if 0 <= 2 * tt + j < t)

    def up(t):
        return sum(1 for tt in range(2 * t) for j in range(4) if (tt + j - 2) % 2 == 0 and 0 <= (tt + j - 2) // 2 < t)

    macs = 0

    def crb(cin, cout, t, proj):
        nonlocal macs
        macs += k5(t) * cin * cout + k5(t) * cout * cout + (t * cin * cout if proj else 0)

    c, t = D, T
    for i, d in enumerate(down):
        crb(c, d, t, True)
        crb(d, d, t, False)
        c = d
        if i < len(down) - 1:
            macs += dn(t) * d * d
            t //= 2
    crb(c, c, t, False)
    crb(c, c, t, False)
    skips = list(down)
    for d in reversed(down[:-1]):
        s = skips.pop()
        crb(c + s, d, t, True)
        crb(d, d, t, False)
        c = d
        macs += up(t) * d * d
        t *= 2
    macs += k5(t) * down[0] * down[0] + t * down[0] * D
    return 2 * macs


# ----------------------------------------------------------------------------------------------------
# clocks during the timed region (profiling recipe's clocks line)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            self.path = tempfile.NamedTemporaryFile(prefix="ldp_clocks_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------
def _oracle_reverse_steps(n_rev: int, reps: int, warm: int, threads: int):
    """Times `reps` samples of `n_rev` consecutive reverse steps (k = 99, 98, ...) of the fp32 CPU oracle at B=1024."""
    import torch
    from latent_diffusion_planning_b200 import params as P
    from oracle import ldp_oracle as O
    torch.set_num_threads(threads)
    p = P.init_params(P.unet_spec(D_OBS, D_OBS), seed=0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B_PLANS, T_PRED, D_OBS, generator=g)
    c = torch.rand(B_PLANS, D_OBS, generator=g) * 2 - 1
    sched = O.ddpm_schedule(N_DIFF)
    times = []
    with torch.no_grad():
        for it in range(warm + reps):
            t0 = time.perf_counter()
            xx = x
            for i in range(n_rev):
                k = N_DIFF - 1 - i
                eps = O.unet_forward(p, xx, k, c, dtype=torch.float32)
                z = torch.randn(xx.shape, generator=g)
                xx = O.ddpm_step(sched, eps, k, xx, z, dtype=torch.float32)
            dt = time.perf_counter() - t0
            if it >= warm:
                times.append(dt)
    return times


def cpu_baseline_leg() -> dict:
    threads = os.cpu_count() or 1
    t1 = _oracle_reverse_steps(1, reps=1, warm=0, threads=threads)[0]          # size the sample: ~15 s of CPU work
    n_rev = max(4, min(N_DIFF, int(15.0 / max(t1, 1e-3))))
    t = _oracle_reverse_steps(n_rev, reps=1, warm=0, threads=threads)[0]
    plans_per_s = B_PLANS / (t / n_rev * N_DIFF)
    return {"value": plans_per_s, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_rev} of the {N_DIFF} DDPM reverse steps at B={B_PLANS} (fp32 PyTorch-CPU oracle, "
                      f"{t:.1f} s), extrapolated linearly to {N_DIFF} steps"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rev = 2
    times = _oracle_reverse_steps(n_rev, reps=args.steps, warm=min(args.warmup, 1), threads=threads)
    t_step = sum(times) / len(times)
    value = B_PLANS / (t_step / n_rev * N_DIFF)
    sample = (f"each step = {n_rev} of the {N_DIFF} DDPM reverse steps at B={B_PLANS}, T={T_PRED}, D={D_OBS} on the fp32 "
              f"PyTorch-CPU oracle (oracle/ldp_oracle.py; the reference's JAX cannot be installed here), "
              f"plans/s extrapolated linearly to {N_DIFF} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "planner denoising loop 100 DDPM steps, B=1024 H=9 (T=8) latent=8x8x4 (D=265)",
                   "B": B_PLANS, "T": T_PRED, "D": D_OBS, "n_diffusion_steps": N_DIFF},
        "denoise_steps_per_sec": value * N_DIFF,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from latent_diffusion_planning_b200 import _native, handles as H, params as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _native.load()
    _native.check(lib.ldp_device_check())

    p = P.init_params(P.unet_spec(D_OBS, D_OBS), seed=0)
    planner = H.Planner(p, D_OBS, D_OBS)
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(B_PLANS, T_PRED, D_OBS, generator=g).pin_memory()
    c_host = (torch.rand(B_PLANS, D_OBS, generator=g) * 2 - 1).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_dev, c_dev = x_host.cuda(), c_host.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    row_offset = rank * B_PLANS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(seed):
        return planner.sample(x_dev, c_dev, seed=seed, row_offset=row_offset, n_steps=N_DIFF, sampler="ddpm", precision="bf16")

    def timed(fn, k):
        """k steps, each bracketed by its own CUDA events on the launching stream, L2 flushed between steps."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        for i in range(k):
            flush.zero_()
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                     # nvidia-smi needs ~1 s to produce its first sample: start it before the warm-up
    t_w = time.perf_counter()
    for i in range(max(args.warmup, 3)):
        one_step(1000 + i)
    torch.cuda.synchronize()
    while sampler and time.perf_counter() - t_w < 1.5:      # keep the GPU under the same load until samples flow
        one_step(2000)
        torch.cuda.synchronize()
    lib.ldp_launch_count_reset()
    ms_total = timed(lambda i: one_step(i), args.steps)
    launches = int(lib.ldp_launch_count())
    clocks = sampler.stop() if sampler else None

    # end to end through the public host-buffer call: pinned host -> device, sample, device -> pinned host
    def e2e_step(i):
        planner.sample_host(x_host, c_host, out_host, seed=i, row_offset=row_offset, n_steps=N_DIFF, sampler="ddpm",
                            precision="bf16")
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PF/s sustained (B200_PROFILING.md)"
        ms_per_step = ms_total / args.steps
        plans_total = B_PLANS * world
        value = plans_total / (ms_per_step / 1e3)
        flops_step = unet_useful_flops(D_OBS, T_PRED) * B_PLANS * N_DIFF      # per rank per bench step
        launches_per_step = launches / args.steps
        avg_launch_us = ms_per_step * 1e3 / launches_per_step
        achieved_tf = flops_step / (ms_per_step / 1e3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        per_op = None
        try:
            ops = planner.profile_step(B_PLANS, T_PRED, reps=20)
            fl = unet_useful_flops(D_OBS, T_PRED) * B_PLANS
            per_op = {"sum_isolated_us": sum(o["us"] for o in ops), "n_kernels": len(ops),
                      "isolated_tflops": fl / (sum(o["us"] for o in ops) * 1e-6) / 1e12,
                      "slowest": sorted(({k: o[k] for k in ("us", "M", "N", "K", "block_n", "epilogue")} for o in ops),
                                        key=lambda o: -o["us"])[:3]}
        except Exception as e:                                              # diagnostics only
            per_op = {"error": str(e)}
        extras = None
        if not args.no_extras:
            try:
                extras = side_metrics(torch, H, P, rank)
            except Exception as e:                                          # side numbers never break the headline line
                extras = {"error": str(e)}
        cpu = cpu_baseline_leg() if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "planner denoising loop 100 DDPM steps, B=1024 H=9 (T=8) latent=8x8x4 (D=265), bf16, per GPU",
                       "B_per_gpu": B_PLANS, "T": T_PRED, "D": D_OBS, "n_diffusion_steps": N_DIFF, "sampler": "ddpm",
                       "noise": "in-kernel Philox4x32-10", "weights": "random init (seed 0), 69.5 M params",
                       "l2": "256 MB buffer written between timed steps (L2 flush); per-step working set (139 MB bf16 "
                             "weights + activations) also exceeds the 126 MB L2",
                       "parallelism": f"dp{world} (independent plans, no collective)"},
            "denoise_steps_per_sec": value * N_DIFF,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "kernel": "tc_gemm_kernel<BN,EPI> (tcgen05 implicit-GEMM conv + fused GN/Mish/FiLM/DDPM epilogues)",
                         "launches_per_step": launches_per_step, "avg_launch_us": avg_launch_us,
                         "algorithmic_flops_per_launch": flops_step / launches_per_step,
                         "peak_source": peak_src + " (of measured)" if "MEASURED" in peak_src else peak_src,
                         "isolated": per_op},
            "e2e": {"value": plans_total / (ms_e2e / args.steps / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": x_host.numel() * 4 + c_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "ms_per_step": ms_e2e / args.steps, "api": "handles.Planner.sample_host (pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "host_cores": os.cpu_count(),
        }
        if extras is not None:
            line["extras"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def side_metrics(torch, H, P, rank):
    """Per-GPU side numbers for the other BASELINE configs (not the headline): VAE-encode img/s at B=4096 (config #3),
    IDM loop and full act() = encode + planner + IDM at B=1024 (rm_lift shapes).  Device-resident inputs, CUDA events."""
    from latent_diffusion_planning_b200.agent import LDPAgent
    out = {}

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    lowdim = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
    shapes = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [LATENT]}
    import numpy as np
    norm = {"obs": {"agentview_image": {"min": 0, "max": 255},
                    "latent_agentview_image": {"min": np.full(LATENT, -10.0, np.float32), "max": np.full(LATENT, 10.0, np.float32)},
                    **{k: {"min": -np.ones(shapes[k][0], np.float32), "max": np.ones(shapes[k][0], np.float32)} for k in lowdim}},
            "actions": {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}}
    agent = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": shapes}, rgb_obs=["latent_agentview_image"], lowdim_obs=lowdim,
                            obs_normalization=norm, vae_feature_dim=LATENT, obs_horizon=1, pred_horizon=T_PRED, action_horizon=4)
    g = torch.Generator().manual_seed(4 + rank)
    img = torch.randint(0, 256, (4096, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8).cuda()
    ms = timeit(lambda: agent.vae.encode(img, lat_min=-10.0, lat_max=10.0, precision="bf16"), 2)
    out["vae_encode_imgs_per_sec"] = 4096 / ms * 1e3
    out["vae_encode_config"] = "stable_vae_model.encode 64x64x3 uint8, B=4096, SD-VAE [128,256,512,512] -> 8x8x4, bf16"
    out["vae_tflops_useful"] = 4096 * 16.92 / ms
    batch = {"obs": {"agentview_image": img[:B_PLANS].reshape(B_PLANS, 1, 64, 64, 3),
                     **{k: torch.rand(B_PLANS, 1, shapes[k][0], generator=g).cuda() * 2 - 1 for k in lowdim}}}
    ms = timeit(lambda: agent.act(batch, 1), 2)
    out["act_plans_per_sec"] = B_PLANS / ms * 1e3
    out["act_config"] = "LDPAgent.act: VAE encode + 100 planner DDPM steps + 100 IDM DDPM steps, B=1024, rm_lift shapes, bf16"
    out["act_ms"] = ms
    # the two remaining pieces of act(), timed alone: IDM reverse loop (B * Ha = 4096 transition rows) and, for sample_viz
    # with viz=True, the VAE decoder over the (Ha + 1) frames of every plan
    idm = agent.idm
    ssp = (torch.rand(B_PLANS * 4, 2 * D_OBS, generator=g) * 2 - 1).cuda()
    a_T = torch.randn(B_PLANS * 4, 7, generator=g).cuda()
    ms = timeit(lambda: idm.sample(ssp, a_T, seed=1, n_steps=N_DIFF, precision="bf16"), 3)
    out["idm_loop_ms"] = ms
    out["idm_config"] = "MLPDiffusion reverse loop, 4096 rows (B=1024 x Ha=4), 2D=530, A=7, 100 DDPM steps, bf16"
    dec = H.VaeDecoder(P.init_params(P.vae_decoder_spec(), seed=7))
    z = torch.randn(1024, 8, 8, 4, generator=g).cuda()
    ms = timeit(lambda: dec.decode(z, precision="bf16"), 2)
    out["vae_decode_frames_per_sec"] = 1024 / ms * 1e3
    out["vae_decode_config"] = "FlaxAutoencoderKL.decode 8x8x4 -> 64x64x3, SD-VAE [128,256,512,512], B=1024 frames, bf16 (plan_viz path)"
    # scope row N1: one LDPAgent.update (planner + IDM losses, gradients, Adam) at the reference's train batch (train_bc.yaml:10)
    g = torch.Generator().manual_seed(5)
    tb = {"obs": {"latent_agentview_image": (torch.randn(256, 9, LATENT, generator=g) * 3).cuda()}, "actions": torch.randn(256, 9, 7, generator=g).cuda()}
    for k in lowdim:
        tb["obs"][k] = (torch.rand(256, 9, shapes[k][0], generator=g) * 2 - 1).cuda()
    agent.data_parallel = False          # side metric of rank 0 alone: no gradient all-reduce here (scripts/train_bench.py has it)
    for i in range(3):
        agent.update(tb, i, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        agent.update(tb, 3 + i, 3 + i)
    e1.record()
    torch.cuda.synchronize()
    out["train_step_ms"] = e0.elapsed_time(e1) / 5
    out["train_samples_per_sec"] = 256 / out["train_step_ms"] * 1e3
    out["train_config"] = "LDPAgent.update (train_bc.py agent=ldp_agent, rm_lift latent_img shapes), batch 256, bf16 tcgen05 contractions, fp32 master weights + Adam"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the side metrics (VAE encode, act())")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
