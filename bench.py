#!/usr/bin/env python
"""bench.py - the planner reverse-diffusion loop (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch: 100 DDPM reverse steps of the planner score network
(ConditionalUnet1D, D = 8*8*4 + 9 = 265) on B = 1024 plans of horizon 9 (T = 8), bf16 tensor-core path, the scheduler
update fused into the last GEMM, in-kernel Philox noise.  Weak scaling: every rank runs its own 1024 plans
(independent units, no data-path collective); `value` = plans of all ranks / max-over-ranks device time.

Besides the headline the same line carries first-class blocks, each timed on every rank and max-reduced over ranks:
  strong    the SAME 1024 plans split over the N ranks (B/N per rank, global-row Philox noise), planner loop, and BASELINE
            config #5 end to end (aloha act(): VAE encode + planner + IDM, B = 512 total, T = 16, 100 DDIM steps, actions
            all-gathered over NCCL) - the split north_star names;
  train_dp  LDPAgent.update with the gradient all-reduce ON: 256 per rank (global 256 N) and fixed global 256, the
            all-reduce timed alone (bus bandwidth) and the same step with the exchange switched off;
  vae       BASELINE config #3: VAE encode at B = 4096 per GPU with its own roofline, host-buffer e2e and CPU baseline;
  parity    the measured abs / max-normalised errors of the benchmarked configurations (profiles/parity_r2.json,
            written from the -m gpu test run of tests/test_bench_config_parity_gpu.py).

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
    n_rev = 8
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
VAE_GFLOP_PER_IMG = 16.92            # 4-block SD-VAE encoder, 64x64x3 -> 8x8x4, useful = nominal
VAE_B = 4096
RM_LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
RM_SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [LATENT]}


class Ctx:
    """torch / torch.distributed plumbing shared by the blocks: barrier + sync, per-step CUDA events on the launching
    stream, L2 flush between timed steps, MAX over ranks."""

    def __init__(self, torch, dist, world, rank, local):
        self.torch, self.dist, self.world, self.rank, self.local = torch, dist, world, rank, local
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, k, warm=0, flush=True):
        """ms per step: `warm` untimed then k timed steps, each bracketed by its own CUDA events, max over ranks of the sum."""
        torch = self.torch
        for i in range(warm):
            fn(-1 - i)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        self.barrier()
        for i in range(k):
            if flush:
                self.flush.zero_()
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / k


def _peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(pk.read_text()) if pk.exists() else {}
    if "bf16_tflops_sustained" in peaks:
        return float(peaks["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)", peaks
    return 1400.0, "fallback 1.4 PF/s sustained (B200_PROFILING.md) (of fallback)", peaks


def _rm_norm(np):
    return {"obs": {"agentview_image": {"min": 0, "max": 255},
                    "latent_agentview_image": {"min": np.full(LATENT, -10.0, np.float32), "max": np.full(LATENT, 10.0, np.float32)},
                    **{k: {"min": -np.ones(RM_SHAPES[k][0], np.float32), "max": np.ones(RM_SHAPES[k][0], np.float32)} for k in RM_LOWDIM}},
            "actions": {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}}


def run_ours(args) -> None:
    import numpy as np
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
    ctx = Ctx(torch, dist, world, rank, local)

    p = P.init_params(P.unet_spec(D_OBS, D_OBS), seed=0)
    planner = H.Planner(p, D_OBS, D_OBS)
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(B_PLANS, T_PRED, D_OBS, generator=g).pin_memory()
    c_host = (torch.rand(B_PLANS, D_OBS, generator=g) * 2 - 1).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_dev, c_dev = x_host.cuda(), c_host.cuda()
    row_offset = rank * B_PLANS

    def one_step(seed):
        return planner.sample(x_dev, c_dev, seed=seed, row_offset=row_offset, n_steps=N_DIFF, sampler="ddpm", precision="bf16")

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
    ms_per_step = ctx.timed(lambda i: one_step(max(i, 0)), args.steps)
    launches = int(lib.ldp_launch_count())
    clocks = sampler.stop() if sampler else None

    # end to end through the public host-buffer call: pinned host -> device, sample, device -> pinned host
    def e2e_step(i):
        planner.sample_host(x_host, c_host, out_host, seed=max(i, 0), row_offset=row_offset, n_steps=N_DIFF, sampler="ddpm",
                            precision="bf16")
    ms_e2e = ctx.timed(e2e_step, args.steps, warm=2)

    blocks = {}
    if not args.no_extras:
        for name, fn in (("strong", strong_block), ("vae", vae_block), ("act", act_block), ("train_dp", train_block)):
            try:
                blocks[name] = fn(ctx, args, planner, x_host, c_host)
            except Exception as e:                                          # a side block never breaks the headline line
                import traceback
                blocks[name] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}
                ctx.barrier()

    if rank == 0:
        peak_tf, peak_src, _ = _peaks()
        plans_total = B_PLANS * world
        value = plans_total / (ms_per_step / 1e3)
        flops_step = unet_useful_flops(D_OBS, T_PRED) * B_PLANS * N_DIFF      # per rank per bench step
        launches_per_step = launches / args.steps
        avg_launch_us = ms_per_step * 1e3 / launches_per_step
        achieved_tf = flops_step / (ms_per_step / 1e3) / 1e12
        traffic, traffic_src = None, None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/traffic.json: " + str(tj.get("source", "ncu --set full capture"))
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
        cpu = cpu_baseline_leg() if world == 1 and not args.no_cpu_baseline else None
        parity = None
        pf = ROOT / "profiles" / "parity_r2.json"
        if pf.exists():
            try:
                pj = json.loads(pf.read_text())
                keys = ("planner_b1024_eps_k50", "planner_b1024_eps_k99", "planner_b1024_eps_k0", "planner_b1024_xprev_k50")
                parity = {"source": "profiles/parity_r2.json (tests/test_bench_config_parity_gpu.py, B=1024 T=8 D=265 bf16 vs float64 oracle)",
                          "abs": max(pj[k]["abs"] for k in keys if k in pj), "rel": max(pj[k]["rel"] for k in keys if k in pj),
                          "tolerance": "north_star 1e-2 gated on rel = abs / max(1, max|ref|); abs is reported, gated at 5e-2",
                          "cases": {k: {"abs": v["abs"], "rel": v["rel"]} for k, v in pj.items()}}
            except Exception:
                parity = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "planner denoising loop 100 DDPM steps, B=1024 H=9 (T=8) latent=8x8x4 (D=265), bf16, per GPU",
                       "B": B_PLANS, "B_per_gpu": B_PLANS, "T": T_PRED, "D": D_OBS, "n_diffusion_steps": N_DIFF, "sampler": "ddpm",
                       "noise": "in-kernel Philox4x32-10", "weights": "random init (seed 0), 69.5 M params",
                       "l2": "256 MB buffer written between timed steps (L2 flush); per-step working set (139 MB bf16 "
                             "weights + activations) also exceeds the 126 MB L2",
                       "parallelism": f"dp{world} (independent plans, no collective)"},
            "denoise_steps_per_sec": value * N_DIFF,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "tc_gemm_kernel_v1<BN,EPI,PAIR> (tcgen05 implicit-GEMM conv + fused GN/Mish/FiLM/DDPM epilogues)",
                         "launches_per_step": launches_per_step, "avg_launch_us": avg_launch_us,
                         "algorithmic_flops_per_launch": flops_step / launches_per_step,
                         "peak_source": peak_src, "isolated": per_op},
            "e2e": {"value": plans_total / (ms_e2e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": x_host.numel() * 4 + c_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "ms_per_step": ms_e2e, "api": "handles.Planner.sample_host (pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "host_cores": os.cpu_count(),
        }
        line.update(blocks)
        if parity is not None:
            line["parity"] = parity
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_block(ctx, args, planner, x_host, c_host):
    """Strong scaling of sampling (the split north_star names; reference call site hands ONE batch: utils/rm_env_utils.py:168-185):
    the same B = 1024 plans sharded over the ranks by rows, no data-path collective, global-row Philox noise so any N draws the
    same numbers.  Then BASELINE config #5: aloha act() end to end, B = 512 total, T = 16, D = 270, A = 14, 100 DDIM steps,
    `LDPAgent.sample_sharded` with the (B, Ha, A) actions all-gathered over NCCL."""
    import numpy as np
    torch = ctx.torch
    from latent_diffusion_planning_b200.agent import LDPAgent, shard_rows
    out = {}
    lo, hi = shard_rows(B_PLANS, ctx.rank, ctx.world)
    g = torch.Generator().manual_seed(1)                         # the SAME global batch on every rank; each takes its rows
    x = torch.randn(B_PLANS, T_PRED, D_OBS, generator=g)[lo:hi].cuda()
    c = (torch.rand(B_PLANS, D_OBS, generator=g) * 2 - 1)[lo:hi].cuda()
    steps = max(3, min(args.steps, 10))
    ms = ctx.timed(lambda i: planner.sample(x, c, seed=max(i, 0), row_offset=lo, n_steps=N_DIFF, sampler="ddpm", precision="bf16"),
                   steps, warm=3)
    fl = unet_useful_flops(D_OBS, T_PRED) * B_PLANS * N_DIFF
    peak_tf, _, _ = _peaks()
    out["planner"] = {"workload": f"planner loop, B={B_PLANS} TOTAL split by rows over {ctx.world} ranks ({hi - lo} per rank), T=8, D=265, 100 DDPM steps, bf16",
                      "ms_per_step": ms, "plans_per_sec": B_PLANS / ms * 1e3, "rows_per_rank": (hi - lo) * T_PRED,
                      "frac_of_n_gpu_tensor_peak": fl / (ms / 1e3) / 1e12 / (peak_tf * ctx.world),
                      "collective": "none (independent plans)",
                      "limit": "M = rows per rank: below ~2048 rows a layer is fewer CTAs than SMs and its fixed cost "
                               "(launch gap + prologue + epilogue) dominates"}
    # config #5
    lat, A, Ha, T5, B5 = 256, 14, 4, 16, 512
    shapes = {"qpos": [14], "latent_wrist64_image": [lat]}
    norm = {"obs": {"wrist64_image": {"min": 0, "max": 255},
                    "latent_wrist64_image": {"min": np.full(lat, -5.5, np.float32), "max": np.full(lat, 5.5, np.float32)},
                    "qpos": {"min": -np.ones(14, np.float32) * 2, "max": np.ones(14, np.float32) * 2}},
            "actions": {"min": -np.ones(A, np.float32) * 1.5, "max": np.ones(A, np.float32) * 1.5}}
    agent = LDPAgent.create(0, None, {"ac_dim": A, "all_shapes": shapes}, rgb_obs=["latent_wrist64_image"], lowdim_obs=["qpos"],
                            obs_normalization=norm, vae_feature_dim=lat, obs_horizon=1, pred_horizon=T5, action_horizon=Ha,
                            sampler="ddim", data_name="aloha_cube")
    g = torch.Generator().manual_seed(9)
    img_host = torch.randint(0, 256, (B5, 1, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8).pin_memory()
    qpos_host = (torch.rand(B5, 1, 14, generator=g) * 2 - 1).pin_memory()
    res = {}

    def act5(i):
        batch = {"obs": {"wrist64_image": img_host.cuda(non_blocking=True), "qpos": qpos_host.cuda(non_blocking=True)}}
        a, _ = agent.sample_sharded(batch, max(i, 0) + 1, gather=True)
        res["a"] = a.to("cpu", non_blocking=True)
    ms = ctx.timed(act5, steps, warm=2)
    out["aloha_act_config5"] = {"workload": f"LDPAgent.sample_sharded: VAE encode + planner + IDM, aloha shapes B={B5} TOTAL over {ctx.world} ranks, "
                                            f"T={T5}, D=270, A={A}, 100 DDIM steps each, bf16; host uint8 frames in, actions all-gathered and read back",
                                "ms_per_act": ms, "plans_per_sec": B5 / ms * 1e3,
                                "collective": "all_gather of (B,Ha,A) actions, %d bytes" % (B5 * Ha * A * 4) if ctx.world > 1 else "none",
                                "actions_finite": bool(torch.isfinite(res["a"]).all()), "actions_shape": list(res["a"].shape)}
    del agent
    return out


def vae_block(ctx, args, planner, x_host, c_host):
    """BASELINE config #3: stable_vae_model.encode over 64x64x3 frames, B = 4096 per GPU (process_sdvae_data path)."""
    torch = ctx.torch
    from latent_diffusion_planning_b200 import handles as H, params as P
    vp = P.init_params(P.vae_encoder_spec(), seed=2)
    vae = H.VaeEncoder(vp)
    g = torch.Generator().manual_seed(4 + ctx.rank)
    img_host = torch.randint(0, 256, (VAE_B, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8).pin_memory()
    img = img_host.cuda()
    lat_host = torch.empty(VAE_B, 8, 8, 4, dtype=torch.float32).pin_memory()
    steps = max(3, min(args.steps, 5))
    lib = vae.lib
    lib.ldp_launch_count_reset()
    ms = ctx.timed(lambda i: vae.encode(img, lat_min=-10.0, lat_max=10.0, precision="bf16"), steps, warm=3)
    launches = int(lib.ldp_launch_count()) / (steps + 3)

    def e2e(i):
        z = vae.encode(img_host.cuda(non_blocking=True), lat_min=-10.0, lat_max=10.0, precision="bf16")
        lat_host.copy_(z, non_blocking=True)
    ms_e2e = ctx.timed(e2e, steps, warm=1)
    peak_tf, peak_src, _ = _peaks()
    ach = VAE_B * VAE_GFLOP_PER_IMG / ms                       # GF / ms = TF/s
    out = {"metric": "vae_encode_imgs_per_sec", "value": VAE_B * ctx.world / ms * 1e3, "unit": "img/s", "ms_per_step": ms,
           "config": {"workload": "stable_vae_model.encode 64x64x3 uint8 agentview images, B=4096 per GPU, SD-VAE [128,256,512,512] "
                                  "-> 8x8x4 latents (+ fused latent normalisation), bf16 tensor-core path, chunks of 592 images",
                      "B_per_gpu": VAE_B, "l2": "256 MB L2 flush between steps; inputs (50 MB) + activations exceed L2"},
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                        "algorithmic_gflop_per_image": VAE_GFLOP_PER_IMG, "launches_per_step": launches, "peak_source": peak_src,
                        "kernel": "tc_gemm_kernel<*, PLAIN, pair, persistent> (3x3 / 1x1 / stride-2 implicit-GEMM convolutions, TMA epilogue)"},
           "e2e": {"value": VAE_B * ctx.world / ms_e2e * 1e3, "unit": "img/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": img_host.numel(), "d2h_bytes_per_step": lat_host.numel() * 4,
                   "api": "handles.VaeEncoder.encode on pinned uint8 host frames, latents copied back"}}
    if ctx.world == 1 and not args.no_cpu_baseline:
        import time as _t
        from oracle import ldp_oracle as O                     # CPU baseline leg (B = 64, fp32)
        torch.set_num_threads(os.cpu_count() or 1)
        sub = img_host[:64].float() / 255 * 2 - 1
        with torch.no_grad():
            O.vae_encode_mean(vp, sub[:8], dtype=torch.float32)
            t0, n_img = _t.perf_counter(), 0
            while _t.perf_counter() - t0 < 12.0:           # batches of 64 for ~12 s of CPU work
                O.vae_encode_mean(vp, sub, dtype=torch.float32)
                n_img += 64
            dt = _t.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_img / dt, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{n_img} images in batches of 64, fp32 PyTorch-CPU oracle ({dt:.1f} s)"}
    vae.close()
    return out


def act_block(ctx, args, planner, x_host, c_host):
    """LDPAgent.act at B = 1024 per GPU (rm_lift shapes): VAE encode + 100 planner steps + 100 IDM steps; IDM loop and VAE
    decoder timed alone.  Weak scaling (every rank its own batch)."""
    import numpy as np
    torch = ctx.torch
    from latent_diffusion_planning_b200 import handles as H, params as P
    from latent_diffusion_planning_b200.agent import LDPAgent
    out = {}
    agent = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": RM_SHAPES}, rgb_obs=["latent_agentview_image"], lowdim_obs=RM_LOWDIM,
                            obs_normalization=_rm_norm(np), vae_feature_dim=LATENT, obs_horizon=1, pred_horizon=T_PRED, action_horizon=4)
    g = torch.Generator().manual_seed(4 + ctx.rank)
    img_host = torch.randint(0, 256, (B_PLANS, 1, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8).pin_memory()
    low_host = {k: (torch.rand(B_PLANS, 1, RM_SHAPES[k][0], generator=g) * 2 - 1).pin_memory() for k in RM_LOWDIM}
    res = {}

    def act(i):
        batch = {"obs": {"agentview_image": img_host.cuda(non_blocking=True), **{k: v.cuda(non_blocking=True) for k, v in low_host.items()}}}
        a, _ = agent.act(batch, max(i, 0) + 1, row_offset=ctx.rank * B_PLANS)
        res["a"] = a.to("cpu", non_blocking=True)
    steps = max(3, min(args.steps, 5))
    ms = ctx.timed(act, steps, warm=2)
    out["act_ms"] = ms
    out["act_plans_per_sec"] = B_PLANS * ctx.world / ms * 1e3
    out["act_config"] = ("LDPAgent.act end to end: pinned host uint8 frames + low-dim -> VAE encode + 100 planner DDPM steps + 100 IDM "
                         "DDPM steps -> actions read back, B=1024 per GPU, rm_lift shapes, bf16")
    idm = agent.idm
    ssp = (torch.rand(B_PLANS * 4, 2 * D_OBS, generator=g) * 2 - 1).cuda()
    a_T = torch.randn(B_PLANS * 4, 7, generator=g).cuda()
    ms = ctx.timed(lambda i: idm.sample(ssp, a_T, seed=1, n_steps=N_DIFF, precision="bf16"), steps, warm=2)
    peak_tf, _, _ = _peaks()
    out["idm_loop_ms"] = ms
    out["idm_roofline_frac"] = 3.153e6 * B_PLANS * 4 * N_DIFF / (ms / 1e3) / 1e12 / peak_tf
    out["idm_config"] = "MLPDiffusion reverse loop, 4096 rows (B=1024 x Ha=4), 2D=530, A=7, 100 DDPM steps, bf16; 3.153 MF useful per row-step"
    dec = H.VaeDecoder(P.init_params(P.vae_decoder_spec(), seed=7))
    z = torch.randn(1024, 8, 8, 4, generator=g).cuda()
    ms = ctx.timed(lambda i: dec.decode(z, precision="bf16"), 3, warm=1)
    out["vae_decode_frames_per_sec"] = 1024 / ms * 1e3
    out["vae_decode_config"] = "FlaxAutoencoderKL.decode 8x8x4 -> 64x64x3, SD-VAE [128,256,512,512], B=1024 frames, bf16 (plan_viz path)"
    dec.close()
    del agent
    return out


def train_block(ctx, args, planner, x_host, c_host):
    """BASELINE config #4: LDPAgent.update (train_bc.py agent=ldp_agent data=cfg/rm_lift/latent_img), bf16 contractions, fp32 master
    weights + Adam, data-parallel gradient all-reduce over NCCL (reference train_bc.py:70-78: the global batch is sharded over devices)."""
    import numpy as np
    torch, dist, world = ctx.torch, ctx.dist, ctx.world
    from latent_diffusion_planning_b200.agent import LDPAgent
    shapes = dict(RM_SHAPES)
    agent = LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": shapes}, rgb_obs=["latent_agentview_image"], lowdim_obs=RM_LOWDIM,
                            obs_normalization={k: v for k, v in _rm_norm(np).items()}, vae_feature_dim=LATENT, obs_horizon=1,
                            pred_horizon=T_PRED, action_horizon=4, vae_params=None)
    agent.vae = None                                      # latents are read from the latent dataset (latent_img config)

    def make_batch(b, seed):
        g = torch.Generator().manual_seed(seed)
        tb = {"obs": {"latent_agentview_image": (torch.randn(b, 9, LATENT, generator=g) * 3).cuda()}, "actions": torch.randn(b, 9, 7, generator=g).cuda()}
        for k in RM_LOWDIM:
            tb["obs"][k] = (torch.rand(b, 9, RM_SHAPES[k][0], generator=g) * 2 - 1).cuda()
        return tb
    out = {}
    step_no = [0]

    def upd(tb):
        def f(i):
            agent.update(tb, step_no[0], step_no[0])
            step_no[0] += 1
        return f
    steps = max(3, min(args.steps, 10))
    n_param = None
    cases = [("weak_256_per_gpu", 256)]
    if world > 1 and 256 % world == 0:
        cases.append(("strong_global_256", 256 // world))
    for label, b_rank in cases:
        tb = make_batch(b_rank, 5 + ctx.rank)
        agent.data_parallel = True
        ms = ctx.timed(upd(tb), steps, warm=4, flush=False)
        entry = {"batch_per_gpu": b_rank, "global_batch": b_rank * world, "ms_per_step": ms, "samples_per_sec": b_rank * world / ms * 1e3,
                 "allreduce": "on" if world > 1 else "n/a (1 rank)"}
        if world > 1:
            agent.data_parallel = False                   # same per-rank work without the exchange: the exposed communication time
            ms_off = ctx.timed(upd(tb), steps, warm=2, flush=False)
            entry["ms_per_step_allreduce_off"] = ms_off
            entry["exposed_comm_ms"] = ms - ms_off
            agent.data_parallel = True
        out[label] = entry
    grads = [agent._train[n].grads for n in ("planner", "idm") if n in agent._train]
    n_bytes = sum(gr.numel() * 4 for gr in grads)
    out["gradient_bytes"] = n_bytes
    if world > 1:
        def ar(i):
            for gr in grads:
                dist.all_reduce(gr, op=dist.ReduceOp.SUM)
        ms_ar = ctx.timed(ar, 10, warm=3, flush=False)
        out["allreduce_alone_ms"] = ms_ar
        out["allreduce_busbw_gbs"] = 2 * (world - 1) / world * n_bytes / (ms_ar / 1e3) / 1e9
        out["allreduce_busbw_reference_gbs"] = "725 GB/s (8 ranks, 1 GiB; B200_PROFILING.md)"
    out["config"] = ("LDPAgent.update: planner (69.5 M) + IDM (1.9 M) losses, backward, Adam; rm_lift latent_img shapes (D=265, T=8, A=7); bf16 "
                     "tcgen05 contractions, fp32 master weights / moments / gradients; one flat fp32 gradient buffer per network all-reduced (NCCL)")
    del agent
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
