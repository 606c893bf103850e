#!/usr/bin/env python
"""Headline benchmark: sampled graphs/sec of the 1000-step reverse-diffusion sampler (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload zinc250k_v2|qm9_s4|...] [--impl reference]

A "step" is ONE reverse-diffusion solver step (for Reverse+Langevin: corrector + predictor = two joint
evaluations of both score networks, noise, norms, updates) over the whole batch; 1000 of them make one
sample, so   value [graphs/s] = batch / (1000 * seconds_per_step).
Every step costs the same (the work is data independent), so K steps measure the full run's rate.

  value     inputs resident in HBM, K steps replayed from one captured CUDA graph, CUDA-event timed; the K-step region is
            repeated --repeats times (default 5) and the MEDIAN is reported (min / max beside it).
  scaling   "strong" by default = BASELINE.json config 3 (the batch of 10 000 SHARDED over the N GPUs); the weak-scaling
            rate (10 000 graphs per GPU) is measured in the same run and reported as `weak_scaling`.
  e2e       the public sampler API (get_pc_sampler / S4_solver closure) called with HOST flags; the timed
            region holds the H2D copy, the on-device prior, K steps and the D2H copy of (x, adj).
  roofline  the dominant kernel (per-kernel CUDA-event timing of eager launches, gdss_prof_*).
  cpu_baseline  the CPU oracle port of the reference algorithm on the host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per-graph FLOPs of one joint evaluation as the reference computes them
WORKLOADS = {
    "zinc250k_v2": dict(ckpt="zinc250k_v2", N=38, F=9, B=10000, sde_x=("VP", .1, 1.), sde_adj=("VE", .2, 1.),
                        predictor="Reverse", corrector="Langevin", snr=.2, scale_eps=.8, evals=2,
                        flop_eval=6.504e6 + 5.849e7, flop_final=2 * 1444 * (46 * 92 + 92 * 92 + 92), flop_final_x=2 * 38 * (41 * 82 + 82 * 82 + 82 * 9), cpu_B=32, cpu_steps=4,
                        transcendentals=(306e3, 646e3),
                        desc="C3 ZINC250k v2 (N=38, F=9, ScoreNetworkX_GMH + ScoreNetworkA) Reverse+Langevin snr .2 scale_eps .8"),
    "qm9_s4": dict(ckpt="qm9", N=9, F=4, B=10000, sde_x=("VE", .1, 1.), sde_adj=("VE", .1, 1.), predictor="S4",
                   corrector="None", snr=.2, scale_eps=.7, evals=1, flop_eval=1.561e5 + 1.260e6, transcendentals=(6.55e3, 16.6e3),
                   flop_final=2 * 81 * (22 * 44 + 44 * 44 + 44), flop_final_x=2 * 9 * (36 * 72 + 72 * 72 + 72 * 4), cpu_B=1000, cpu_steps=8,
                   desc="C2 QM9 (N=9, F=4, ScoreNetworkX + ScoreNetworkA) S4 snr .2 scale_eps .7"),
    "community_small": dict(ckpt="community_small", N=20, F=10, B=100, sde_x=("VP", .1, 1.), sde_adj=("VP", .1, 1.),
                            predictor="Euler", corrector="Langevin", snr=.05, scale_eps=.7, evals=2,
                            flop_eval=2.953e6 + 1.663e7, flop_final=2 * 400 * (38 * 76 + 76 * 76 + 76), cpu_B=100,
                            transcendentals=(59.5e3, 102.9e3),
                            cpu_steps=8, desc="C1 community_small (N=20) Euler+Langevin"),
    "enzymes": dict(ckpt="enzymes", N=125, F=10, B=64, sde_x=("VP", .1, 1.), sde_adj=("VE", .2, 1.), predictor="S4",
                    corrector="None", snr=.15, scale_eps=.7, evals=1, flop_eval=5.030e7 + 8.747e8, transcendentals=(3.17e6, 5.22e6),
                    flop_final=2 * 15625 * (54 * 108 + 108 * 108 + 108), cpu_B=8, cpu_steps=3, ema=True,
                    desc="C4 ENZYMES (N=125) S4"),
    "grid": dict(ckpt="grid", N=361, F=5, B=8, sde_x=("VP", .1, 1.), sde_adj=("VP", .2, .8), predictor="Reverse",
                 corrector="Langevin", snr=.1, scale_eps=.7, evals=2, flop_eval=1.639e8 + 7.113e9, transcendentals=(26.2e6, 43.0e6),
                 flop_final=2 * 130321 * (54 * 108 + 108 * 108 + 108), cpu_B=1, cpu_steps=2, ema=True,
                 desc="C4 grid (N=361) Reverse+Langevin"),
}


# BASELINE config C5: one optimisation step of trainer.py:57-79 on a ZINC250k batch (config/zinc250k.yaml: batch 1024, Adam
# lr 5e-3, weight decay 1e-4, grad_norm 1, EMA 0.999), data parallel over the ranks
TRAIN_WORKLOADS = {
    "zinc_train": dict(ckpt="zinc250k_v2", N=38, F=9, B=1024, sde_x=("VP", .1, 1.), sde_adj=("VE", .2, 1.), lr=0.005,
                       weight_decay=0.0001, grad_norm=1.0, ema_decay=0.999, cpu_B=16,
                       # forward + backward ~ 3x the forward GEMM count of one joint evaluation
                       flop_step=3 * (6.504e6 + 5.849e7),
                       desc="C5 ZINC250k DSM training step (ScoreNetworkX_GMH + ScoreNetworkA, B=1024, Adam + clip + EMA)"),
}


def synthetic_molecules(wl, B, seed=42):
    """ZINC-like synthetic batch (the real ZINC250k csv is not in the reference tree): node counts from the SURVEY 8d recipe,
    one-hot atom types, symmetric integer bond orders (0..3) with ~1.1 bonds per atom; every atom has at least one bond, so
    node_flags(adj) (losses.py:44) recovers the node counts."""
    N, F = wl["N"], wl["F"]
    rs = np.random.RandomState(seed)
    counts = np.clip(np.round(rs.normal(23.2, 4.5, B)), 6, N).astype(int)
    x = np.zeros((B, N, F), np.float32)
    adj = np.zeros((B, N, N), np.float32)
    for b, n in enumerate(counts):
        x[b, np.arange(n), rs.randint(0, F, n)] = 1.0
        for i in range(1, n):                                   # a random spanning tree, then a few ring-closing bonds
            j = rs.randint(0, i)
            adj[b, i, j] = adj[b, j, i] = rs.choice([1, 2, 3], p=[.85, .13, .02])
        for _ in range(max(1, n // 8)):
            i, j = rs.randint(0, n, 2)
            if i != j and adj[b, i, j] == 0:
                adj[b, i, j] = adj[b, j, i] = 1.0
    return torch.from_numpy(x), torch.from_numpy(adj), counts


def cpu_train_rate(wl, weights, B, steps, device="cpu"):
    """steps/s of the oracle's DSM losses under torch.autograd + torch.optim.Adam + clip + EMA (trainer.py:57-79 restated with
    plain torch ops) on the host cores (or as eager CUDA ops on the GPU)."""
    from oracle import gdss_oracle as O
    sdx, px, sda, pa = weights
    torch.set_num_threads(os.cpu_count() or 1)
    x, adj, _ = synthetic_molecules(wl, B)
    x, adj = x.to(device), adj.to(device)
    px_ = {k: torch.nn.Parameter(v.clone().float().to(device)) for k, v in sdx.items()}
    pa_ = {k: torch.nn.Parameter(v.clone().float().to(device)) for k, v in sda.items()}
    ox = torch.optim.Adam(list(px_.values()), lr=wl["lr"], weight_decay=wl["weight_decay"])
    oa = torch.optim.Adam(list(pa_.values()), lr=wl["lr"], weight_decay=wl["weight_decay"])
    sx, sa = O.SDESpec(*wl["sde_x"], 1000), O.SDESpec(*wl["sde_adj"], 1000)
    shadows = [[p.detach().clone() for p in d.values()] for d in (px_, pa_)]
    torch.manual_seed(0)
    times = []
    for it in range(steps + 1):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ox.zero_grad(); oa.zero_grad()
        t = torch.rand(B, device=device) * (1 - 1e-5) + 1e-5
        lx, la = O.dsm_loss(sx, sa, px_, ck_params(weights, 0), pa_, ck_params(weights, 1), x, adj, t, torch.randn_like(x),
                            torch.randn_like(adj), reduce_mean=False)
        lx.backward(); la.backward()
        torch.nn.utils.clip_grad_norm_(list(px_.values()), wl["grad_norm"])
        torch.nn.utils.clip_grad_norm_(list(pa_.values()), wl["grad_norm"])
        ox.step(); oa.step()
        with torch.no_grad():
            for sh, d in zip(shadows, (px_, pa_)):
                for s_, p_ in zip(sh, d.values()):
                    s_.sub_((1 - wl["ema_decay"]) * (s_ - p_))
        float(lx); float(la)
        if device != "cpu":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sec = float(np.median(times[1:]))
    return 1.0 / sec, sec, torch.get_num_threads()


def ck_params(weights, which):
    return weights[1] if which == 0 else weights[3]


def run_train_bench(args):
    """bench.py --workload zinc_train: `value` = optimisation steps / s of the whole job (global batch 1024 split over the
    ranks: the reference's DataParallel splits one batch the same way), device-timed with the batch resident; e2e = the same
    through DSMTrainStep.step with HOST (pinned) batches and a D2H read of both losses every step."""
    import torch.distributed as dist
    wl = TRAIN_WORKLOADS[args.workload]
    weights = load_weights(wl)
    rank = int(os.environ.get("RANK", "0"))
    B = args.batch or wl["B"]
    if args.impl == "reference":
        if rank != 0:
            return
        cb = args.cpu_batch or wl["cpu_B"]
        v, sec, cores = cpu_train_rate(wl, weights, cb, max(2, min(args.steps, 3)))
        sample = f"{max(2, min(args.steps, 3))} timed training steps (after 1) on B={cb} graphs, {sec:.2f} s/step; graphs/s = {cb / sec:.2f}"
        print(json.dumps({"impl": "reference", "metric": "DSM training steps/sec (ZINC250k, batch 1024)", "value": v * cb / B,
                          "unit": "steps/s (scaled to the batch of 1024 from the sample's graphs/s)", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * B / cb, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic ZINC-like molecules",
                          "config": {"workload": wl["desc"], "batch": B},
                          "cpu_baseline": {"value": v * cb / B, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v * cb / B, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    from gdss_b200 import _lib
    from gdss_b200.dist import init_from_env
    from gdss_b200.trainer import DSMTrainStep
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the GDSS kernels have no CPU fallback")
    shard, dev = init_from_env()
    world, rank = shard.world, shard.rank
    L = _lib.lib()
    sdx, px, sda, pa = weights
    x, adj, counts = synthetic_molecules(wl, B)
    lo, hi = shard.bounds(B)
    K, W, R = args.steps, max(args.warmup, 3), max(args.repeats, 1)
    tr = DSMTrainStep(sdx, sda, wl["N"], sde_from(wl["sde_x"], 1000), sde_from(wl["sde_adj"], 1000), lr=wl["lr"],
                      weight_decay=wl["weight_decay"], grad_norm=wl["grad_norm"], ema=wl["ema_decay"], reduce_mean=False, device=str(dev),
                      shard=shard if world > 1 else None, use_graph=not args.train_eager)
    xd, ad = x[lo:hi].to(dev), adj[lo:hi].to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def mx(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    torch.manual_seed(1234 + rank)                       # every rank draws its own t / noise for its own shard
    for _ in range(W):
        out = tr.step(xd, ad)
    barrier()
    clocks = ClockSampler(dev.index if dev.index is not None else 0, world) if rank == 0 else None
    times = []
    before = L.gdss_launch_count()
    for _ in range(R):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            out = tr.step(xd, ad)
        ev1.record()
        barrier()
        times.append(mx(ev0.elapsed_time(ev1)))
    launches = (L.gdss_launch_count() - before) // R
    used_graph = tr.use_graph
    if used_graph:                                       # replayed launches are not counted by the library: K x the captured step
        launches = K * tr.captured_launches
    clk = clocks.stop() if clocks is not None else {}
    ms = float(np.median(times)) / K
    value = 1000.0 / ms
    losses = [float(out[0]), float(out[1])]
    # e2e: host batches in, losses out, every step
    hx, ha = x[lo:hi].pin_memory(), adj[lo:hi].pin_memory()
    e2e_t = []
    for _ in range(min(R, 3)):
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            o = tr.step(hx.to(dev, non_blocking=True), ha.to(dev, non_blocking=True))
            lx_host, la_host = float(o[0]), float(o[1])
        torch.cuda.synchronize(dev)
        e2e_t.append(mx(time.perf_counter() - t0))
    e2e = {"value": K / float(np.median(e2e_t)), "unit": "steps/s", "h2d_bytes_per_step": (hx.numel() + ha.numel()) * 4,
           "d2h_bytes_per_step": 8, "call": "DSMTrainStep.step(x_host, adj_host) + float(loss_x), float(loss_adj); bytes per rank"}
    tr.close()                                           # captured graphs go before the communicator they reference
    if rank != 0:
        shard.close()
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    ach = value * B * wl["flop_step"] / 1e12
    fp32_peak = 148 * 128 * 2 * (clk.get("sm_mhz") or peaks["sm_max_mhz"]) * 1e6 / 1e12
    roofline = {"bound": "tensor", "kernel": "whole training step (HBM-staged fp32 FFMA kernels: tr_linear / tr_wgrad / tr_gcn_* / tr_maps_*)",
                "achieved": ach / world, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / world / peaks["bf16_tflops_sustained"], "traffic": None, "peak_source": peak_src,
                "frac_of_fp32_ffma_peak": ach / world / fp32_peak,
                "note": "algorithmic FLOPs = 3 x the reference's dense forward GEMM count of both networks per graph (forward + data "
                        "gradient + weight gradient), at padded N with no credit for symmetry / flag skipping; per GPU.  HBM-staged "
                        "multi-kernel formulation on the FP32 pipe: per-edge MLPs and their gradients on one compact row per valid "
                        "upper-triangle pixel, per-graph kernels bounded by the node count (DESIGN 7b)"}
    cpu, torch_gpu = None, None
    if not args.no_cpu_baseline and world == 1:
        cb = args.cpu_batch or wl["cpu_B"]
        v, sec, cores = cpu_train_rate(wl, weights, cb, 2)
        cpu = {"value": v * cb / B, "unit": "steps/s (batch 1024, scaled from the sample's graphs/s)", "cores": cores, "kind": "port",
               "sample": f"2 timed training steps (after 1) on B={cb} graphs, {sec:.2f} s/step"}
    if not args.no_torch_baseline and world == 1:
        try:
            del tr
            torch.cuda.empty_cache()
            v, sec, _ = cpu_train_rate(wl, weights, B, 3, device=str(dev))
            torch_gpu = {"value": v, "unit": "steps/s", "kind": "port (oracle losses under torch.autograd + torch.optim.Adam, eager CUDA fp32, TF32 off)",
                         "sample": f"3 timed steps (after 1) on B={B} on the same B200, {sec * 1e3:.1f} ms/step", "speedup_of_value": value / v}
        except Exception as exc:
            torch_gpu = {"unavailable": repr(exc)[:200]}
    line = {"metric": "DSM training steps/sec (ZINC250k, batch 1024)", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic ZINC-like molecules (node counts: SURVEY 8d recipe), shipped checkpoint as the initial weights",
            "config": {"workload": wl["desc"], "batch": B, "batch_per_gpu": hi - lo, "graphs_per_s": value * B,
                       "sharding": f"data parallel over {world} ranks, one NCCL all-reduce of each network's flat gradient per step" if world > 1 else "none",
                       "l2": "activation arena >> L2 (several GB at B=1024)",
                       "launch": "one CUDA graph per step, replayed" if used_graph else "eager launches"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches) * world, "roofline": roofline, "cpu_baseline": cpu,
            "torch_eager_same_gpu": torch_gpu, "losses_last_step": losses,
            "repeats": {"regions": R, "steps_per_region": K, "ms_per_step_min": min(times) / K, "ms_per_step_max": max(times) / K}}
    print(json.dumps(line), flush=True)
    shard.close()
    if world > 1:
        dist.destroy_process_group()


def make_flags(wl, B, mode):
    """SURVEY.md 8d synthetic node-count histograms (RandomState(42)); 'full' = flags == 1 (worst case)."""
    N = wl["N"]
    rs = np.random.RandomState(42)
    if mode == "full":
        counts = np.full(B, N)
    elif wl["ckpt"].startswith("zinc"):
        counts = np.clip(np.round(rs.normal(23.2, 4.5, B)), 6, 38).astype(int)
    elif wl["ckpt"] == "qm9":
        counts = rs.choice([9, 8, 7, 6, 5], size=B, p=[.83, .14, .025, .004, .001])
    else:
        # node counts of the training graphs (data/<name>.pkl, committed as tests/golden/node_counts.json), drawn like
        # init_flags does (utils/graph_utils.py:66-74: a uniformly drawn training graph)
        hist = json.load(open(os.path.join(ROOT, "tests", "golden", "node_counts.json")))[wl["ckpt"]]["node_counts"]
        counts = np.minimum(np.asarray(hist)[rs.randint(0, len(hist), size=B)], N)
    flags = torch.zeros(B, N)
    for b, c in enumerate(counts):
        flags[b, :c] = 1.0
    return flags, counts


def load_weights(wl):
    from gdss_b200.ckpt import load_ckpt, sampling_weights      # plain torch.load of tests/golden/ckpt (+ EMA, sampler.py:47-52)
    return sampling_weights(load_ckpt(wl["ckpt"]), use_ema=bool(wl.get("ema")))


def sde_from(spec, scales):
    from gdss_b200.sde import VPSDE, VESDE, subVPSDE
    kind, lo, hi = spec
    return {"VP": VPSDE, "VE": VESDE, "subVP": subVPSDE}[kind](lo, hi, scales)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region by ONE process (rank 0) for all
    the GPUs of the job: a poller per rank (8 concurrent nvidia-smi loops on an 8-GPU box) perturbs the run it watches."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, primary, n_gpus=1):
        self.rows, self.proc, self.primary, self.n_gpus = [], None, int(primary), int(n_gpus)
        try:
            sel = [f"--id={primary}"] if n_gpus == 1 else []
            self.proc = subprocess.Popen(["nvidia-smi", *sel, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        num = lambda v: v.replace(".", "").isdigit()
        rows = [r for r in self.rows if len(r) >= 8 and r[0].isdigit() and num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summary(sel):
            sm = [float(r[1]) for r in sel]
            mx = [float(r[2]) for r in sel if num(r[2])]
            pw = [float(r[3]) for r in sel if num(r[3])]
            reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in sel)]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}

        mine = [r for r in rows if int(r[0]) == self.primary] or rows
        out = summary(mine)
        if self.n_gpus > 1:
            per = [summary([r for r in rows if int(r[0]) == g]) for g in sorted({int(r[0]) for r in rows})][: self.n_gpus]
            meds = [p["sm_mhz"] for p in per if p["sm_mhz"] is not None]
            out["sm_mhz_min_over_gpus"] = min(meds) if meds else None
            out["power_w_max_over_gpus"] = max([p["power_w_max"] or 0.0 for p in per]) if per else None
            out["reasons_any_gpu"] = sorted({n for p in per for n in p["reasons"]})
            out["gpus_sampled"] = len(per)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


def mufu_ceiling(wl, graphs_per_s, sm_mhz, n_gpus=1, sms=148):
    """BASELINE.md 3's secondary ceiling: the tanh / ELU elements of the reference's dense formulation (per joint evaluation
    and graph, at padded N) against the special-function pipe, 16 results per clock and SM.  One element is counted as ONE
    MUFU operation, as BASELINE.md counts them (the kernels spend ex2 + rcp on a tanh, and skip padded nodes and the
    symmetric half, so this is a dense-equivalent rate like `roofline.achieved`)."""
    tanh, elu = wl["transcendentals"]
    per_graph = (tanh + elu) * wl["evals"] * 1000.0                     # per sampled graph (1000 solver steps)
    peak = sms * 16 * float(sm_mhz) * 1e6 * n_gpus
    return {"pipe": "MUFU (16 / clk / SM)", "elements_per_graph_step": (tanh + elu) * wl["evals"], "peak_elements_per_s": peak,
            "ceiling_graphs_per_s": peak / per_graph, "achieved_elements_per_s": graphs_per_s * per_graph,
            "frac": graphs_per_s * per_graph / peak,
            "note": "dense-equivalent element count of the reference formulation (BASELINE.md 3), one MUFU operation per element"}


def measured_traffic(workload):
    """DRAM bytes per graph-evaluation of each kernel class, from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the launches of one evaluation, divided by
    the graphs of that run)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(p):
        return json.load(open(p)).get(workload, {})
    return {}


class TimedTrace(list):
    def __init__(self, sync=False):
        super().__init__()
        self.times, self.sync = [], sync

    def append(self, rec):
        if self.sync:
            torch.cuda.synchronize()
        self.times.append(time.perf_counter())


def cpu_oracle_rate(wl, weights, B, warm, steps, flag_mode, device="cpu"):
    """graphs/s of the oracle port (reference algorithm as plain PyTorch fp32 ops) on a bounded sample:
    `steps` timed solver steps (after `warm`) on B graphs -- on the host cores (device='cpu', all threads)
    or as PyTorch eager CUDA ops on the same GPU (device='cuda', TF32 off = the reference's default)."""
    from oracle import gdss_oracle as O
    sdx, px, sda, pa = weights
    N, F = wl["N"], wl["F"]
    torch.set_num_threads(os.cpu_count() or 1)
    flags, _ = make_flags(wl, B, flag_mode)
    draw = O.global_rng_draw
    if device != "cpu":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        sdx = {k: v.to(device) for k, v in sdx.items()}
        sda = {k: v.to(device) for k, v in sda.items()}
        flags = flags.to(device)
        draw = O.device_rng_draw(device)
    sx, sa = O.SDESpec(*wl["sde_x"], 1000), O.SDESpec(*wl["sde_adj"], 1000)
    tr = TimedTrace(sync=device != "cpu")
    torch.manual_seed(0)
    t0 = time.perf_counter()
    with torch.no_grad():
        if wl["predictor"] == "S4":
            O.s4_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw, snr=wl["snr"],
                        scale_eps=wl["scale_eps"], eps=1e-4, trace=tr, max_iters=warm + steps)
        else:
            O.pc_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw,
                        predictor=wl["predictor"], corrector=wl["corrector"], snr=wl["snr"], scale_eps=wl["scale_eps"],
                        eps=1e-4, trace=tr, max_iters=warm + steps)
    start = tr.times[warm - 1] if warm > 0 else t0
    sec_per_step = (tr.times[-1] - start) / steps
    return B / (1000.0 * sec_per_step), sec_per_step, torch.get_num_threads()


def run_reference_arm(args, wl, weights):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.cpu_batch or wl["cpu_B"]
    val, sps, cores = cpu_oracle_rate(wl, weights, B, args.warmup, args.steps, args.flags)
    sample = f"{args.steps} timed solver steps (after {args.warmup}) of 1000 on B={B} graphs, shipped checkpoint weights"
    line = {"impl": "reference", "metric": "sampled graphs/sec (1000-step reverse SDE)", "value": val, "unit": "graphs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic flags",
            "config": workload_config(args, wl, B),
            "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is pure Python/PyTorch and /root/reference is absent on the GPU box: this arm times the "
                    "CPU oracle (oracle/gdss_oracle.py, pinned bit-close to the unmodified reference) on the host cores"}
    print(json.dumps(line), flush=True)


def workload_config(args, wl, B):
    world = max(1, int(os.environ.get("WORLD_SIZE", "1")))
    return {"workload": wl["desc"], "batch": B, "batch_per_gpu": B // world if args.impl != "reference" else B,
            "sharding": "none (CPU arm: rank 0 alone runs it)" if args.impl == "reference" else
                        f"{world} contiguous shards of independent graphs, one 4-float all-reduce per solver phase" if world > 1 else "none",
            "flags": args.flags, "solver_steps_per_sample": 1000,
            "step": "one reverse-diffusion solver step over the whole batch", "l2": "inputs larger than L2 (per-step working "
            "set = edge-channel stacks >> 126 MB at B=10000)" if B * wl["N"] ** 2 * 46 * 4 > 126e6 else "working set fits L2",
            "weights": "shipped checkpoint (tests/golden/ckpt); dead weight blocks (|w| < 1e-12, the checkpoint's weight-decayed "
                       "sub-networks) are not evaluated -- `extra.no_dead_chain_elimination` is the same run with every chain evaluated",
            "noise": "in-kernel Philox4x32-10, seed 42"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="zinc250k_v2", choices=list(WORKLOADS) + list(TRAIN_WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-batch", type=int, default=0)
    ap.add_argument("--flags", default="ragged", choices=["ragged", "full"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE.json config 3): the BASELINE batch in total, sharded over the ranks; weak: the "
                         "BASELINE batch per GPU.  At N > 1 the other mode is measured too and reported as an extra key")
    ap.add_argument("--repeats", type=int, default=5, help="timed K-step regions; the median is reported")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra lines (flags=full, no-prune, C1 / C2 / C4 workloads)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"],
                    help="fp32 (default, the parity build: 3xTF32 split, <= 1e-4) or tf32 (libgdss_b200_tf32.so: single-pass "
                         "TF32, <= 3e-2; a reduced-precision line, never the headline)")
    ap.add_argument("--train-eager", action="store_true", help="zinc_train: launch every step eagerly instead of replaying one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--torch-batch", type=int, default=10000)
    ap.add_argument("--burn-in", type=int, default=20, help="untimed solver steps before the warm-up steps")
    args = ap.parse_args()
    if args.precision == "tf32":
        os.environ["GDSS_PRECISION"] = "tf32"           # read by gdss_b200._lib at its first use
    if args.workload in TRAIN_WORKLOADS:
        return run_train_bench(args)
    wl = WORKLOADS[args.workload]
    weights = load_weights(wl)
    if args.impl == "reference":
        return run_reference_arm(args, wl, weights)

    from gdss_b200 import _lib
    from gdss_b200.dist import init_from_env
    from gdss_b200.engine import Engine, make_sampler_desc
    from gdss_b200.schedule import build_table
    from gdss_b200.solver import get_pc_sampler, S4_solver
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the GDSS kernels have no CPU fallback")
    shard, dev = init_from_env()
    world, rank = shard.world, shard.rank
    L = _lib.lib()
    K, W, R = args.steps, max(args.warmup, 0), max(args.repeats, 1)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def setup(wl_, weights_, B_total, flag_mode, env=None):
        """Engine + schedule + this rank's shard of the flags for a workload at a global batch."""
        sdx_, _, sda_, _ = weights_
        flags_, counts_ = make_flags(wl_, B_total, flag_mode)
        lo_, hi_ = shard.bounds(B_total)
        sx_, sa_ = sde_from(wl_["sde_x"], 1000), sde_from(wl_["sde_adj"], 1000)
        table_ = build_table(sx_, sa_, wl_["predictor"], wl_["corrector"], wl_["snr"], wl_["scale_eps"], 1e-4)
        while table_.shape[0] < W + K + 1:              # K beyond one sample: keep stepping with the same rows
            table_ = np.concatenate([table_, table_[-1000:]])
        sdesc_ = make_sampler_desc(sx_, sa_, wl_["predictor"], wl_["corrector"], wl_["snr"], wl_["scale_eps"], 1e-4, 1, False, True)
        old = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update(env or {})
        try:
            eng_ = Engine(sdx_, sda_, wl_["N"], dev)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        kw_ = dict(seed=42, graph_offset=lo_, global_B=B_total, comm=shard.comm, use_graph=True)
        return dict(eng=eng_, table=table_, sdesc=sdesc_, flags=flags_, counts=counts_, lo=lo_, hi=hi_, fl=flags_[lo_:hi_].to(dev),
                    kw=kw_, sx=sx_, sa=sa_, B=B_total)

    def timed(S, burn, repeats):
        """burn untimed steps, W warm-up steps, then `repeats` regions of EXACTLY K steps each: barrier + synchronize on both
        sides, CUDA events on the launching stream, max over the ranks per region -> list of ms per region."""
        eng_ = S["eng"]
        st = eng_.sample(S["sdesc"], S["table"], S["fl"], n_iters=burn, **S["kw"])
        barrier()
        st = eng_.sample(S["sdesc"], S["table"], S["fl"], n_iters=max(W, 1), **S["kw"])
        barrier()
        out, own = [], []
        before_ = L.gdss_launch_count()
        for _ in range(repeats):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            st = eng_.sample(S["sdesc"], S["table"], S["fl"], x=st[0], adj=st[1], first_step=max(W, 1), n_iters=K, **S["kw"])
            ev1.record()
            barrier()
            own.append(ev0.elapsed_time(ev1))
            out.append(max_over_ranks(own[-1]))
        return out, own, st, (L.gdss_launch_count() - before_) // repeats

    def rate(S, ms):
        return S["B"] / (1000.0 * (ms / K) * 1e-3)

    sdx, px, sda, pa = weights
    N, F = wl["N"], wl["F"]
    B = (args.batch or wl["B"]) * (world if args.scaling == "weak" else 1)
    S = setup(wl, weights, B, args.flags)
    lo, hi, counts, flags = S["lo"], S["hi"], S["counts"], S["flags"]
    eng, table, sdesc, fl_local, sx, sa = S["eng"], S["table"], S["sdesc"], S["fl"], S["sx"], S["sa"]
    clocks = ClockSampler(dev.index if dev.index is not None else 0, world) if rank == 0 else None
    times, own, (x, adj, xm, am), launches = timed(S, args.burn_in, R)
    clk = clocks.stop() if clocks is not None else {}
    ms_total = float(np.median(times))
    if world > 1:
        lo_t = torch.tensor([float(np.median(own))], device=dev)
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
        clk["rank_ms_min_max"] = [float(lo_t[0]), ms_total]
    ms_per_step = ms_total / K
    value = rate(S, ms_total)
    repeats_info = {"regions": R, "steps_per_region": K, "ms_per_step_median": ms_per_step, "ms_per_step_min": min(times) / K,
                    "ms_per_step_max": max(times) / K, "value_min": rate(S, max(times)), "value_max": rate(S, min(times))}
    finite = bool(torch.isfinite(am).all())

    # ---- e2e: public API with host buffers
    get = S4_solver if wl["predictor"] == "S4" else get_pc_sampler
    fn = get(sx, sa, (B, N, F), (B, N, N), predictor=wl["predictor"], corrector=wl["corrector"], snr=wl["snr"],
             scale_eps=wl["scale_eps"], n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-4,
             device=str(dev), shard=shard if world > 1 else None, max_iters=K)
    host_flags = flags.pin_memory()
    out_x = torch.empty(hi - lo, N, F).pin_memory()     # every rank reads back its own shard of the gathered result
    out_a = torch.empty(hi - lo, N, N).pin_memory()
    torch.manual_seed(42)
    fn(sdx, sda, host_flags)                            # warm (engine cache, allocator)
    e2e_t = []
    for _ in range(min(R, 3)):
        barrier()
        t0 = time.perf_counter()
        ex, ea, _ = fn(sdx, sda, host_flags)
        out_x.copy_(ex[lo:hi], non_blocking=True)
        out_a.copy_(ea[lo:hi], non_blocking=True)
        torch.cuda.synchronize(dev)
        e2e_t.append(max_over_ranks(time.perf_counter() - t0))
    e2e_val = B * K / (1000.0 * float(np.median(e2e_t)))
    e2e = {"value": e2e_val, "unit": "graphs/s", "h2d_bytes_per_step": host_flags.numel() * 4 / K,
           "d2h_bytes_per_step": (out_x.numel() + out_a.numel()) * 4 / K, "calls": len(e2e_t),
           "value_min": B * K / (1000.0 * max(e2e_t)), "value_max": B * K / (1000.0 * min(e2e_t)),
           "call": f"sampler(model_x, model_adj, host_flags) for {K} solver steps (+ the final NCCL gather at N > 1) + D2H of "
                   f"the rank's shard of (x, adj); median of {len(e2e_t)} calls; bytes are per rank and call / {K}"}
    del ex, ea

    # ---- the other scaling mode at N > 1 (every rank takes part), extra lines at N = 1
    other_scaling = None
    if world > 1 and not args.no_extra:
        B2 = (args.batch or wl["B"]) * (world if args.scaling == "strong" else 1)
        S2 = setup(wl, weights, B2, args.flags)
        t2, _, _, _ = timed(S2, 10, 3)
        m2 = float(np.median(t2))
        other_scaling = {"scaling": "weak" if args.scaling == "strong" else "strong", "batch": B2, "value": rate(S2, m2),
                         "unit": "graphs/s", "ms_per_step": m2 / K, "regions": 3}
        del S2
    if rank != 0:
        shard.close()
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_src = measured_peaks()
    pruned = int(L.gdss_ctx_query(eng.ctx, b"pruned"))
    # ---- per-kernel timing of eager launches (1 GPU view) -> roofline of the dominant kernel
    roofline, roofline_final, kernel_ms = None, None, None
    if not args.no_profile:
        nb = hi - lo
        L.gdss_prof_begin(torch.cuda.current_stream(dev).cuda_stream)
        # rank 0 alone (the other ranks have returned): its own shard as a stand-alone batch, no collective
        eng.sample(sdesc, table, fl_local, x=x, adj=adj, first_step=1, n_iters=2, seed=42, graph_offset=lo, global_B=nb,
                   comm=None, use_graph=False)
        import ctypes as C
        msb = (C.c_float * 16)()
        cnt = (C.c_int64 * 16)()
        _lib.check(L.gdss_prof_end(msb, cnt, 16), "gdss_prof_end")
        kernel_ms = {L.gdss_prof_name(i).decode(): {"ms_per_step": msb[i] / 2, "launches_per_step": cnt[i] / 2}
                     for i in range(15) if cnt[i]}
        tot = sum(v["ms_per_step"] for v in kernel_ms.values())
        for v in kernel_ms.values():
            v["share"] = v["ms_per_step"] / tot
        peak_tc = peaks["bf16_tflops_sustained"]
        fp32_peak = 148 * 128 * 2 * (clk["sm_mhz"] or peaks["sm_max_mhz"]) * 1e6 / 1e12
        traffic = measured_traffic(args.workload)
        evals = wl["evals"]
        exec_frac = float(np.mean(counts[lo:hi] * (counts[lo:hi] - 1) / 2.0) / (N * N))

        def roof(name, flops_per_graph_eval, note, pipe_peak, pipe_name):
            k = kernel_ms.get(name)
            if not k:
                return None
            ms_eval = k["ms_per_step"] / evals          # all launches of this kernel class in one joint evaluation
            ach = nb * flops_per_graph_eval / (ms_eval * 1e-3) / 1e12
            t = traffic.get(name)
            return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak_tc, "unit": "TFLOP/s", "frac": ach / peak_tc,
                    "traffic": (t["dram_bytes_per_graph_eval"] * nb if t else None),
                    "traffic_source": (t["source"] if t else None),
                    "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                    "ms_per_evaluation": ms_eval, "launches_per_evaluation": k["launches_per_step"] / evals,
                    "share_of_step": k["share"], "pipe": pipe_name, "pipe_peak_tflops": pipe_peak,
                    "frac_of_pipe_peak": ach / pipe_peak,
                    # share of the dense N x N pixel work the kernels really execute (strict upper triangle of the n_b valid
                    # nodes): `achieved` counts the reference's padded dense FLOPs, so frac_of_pipe_peak can exceed 1
                    "executed_pixel_fraction": exec_frac, "frac_of_pipe_peak_executed_work": ach * exec_frac / pipe_peak,
                    "note": note}

        # the dominant kernel first: the fused per-graph layer kernel (everything except the two final MLPs)
        flop_final_x = wl.get("flop_final_x", 0.0)
        mma_peak = 148 * 512 * 2 * (clk["sm_mhz"] or peaks["sm_max_mhz"]) * 1e6 / 1e12 / 3.0
        roofline = roof("fused_layers_kernel", wl["flop_eval"] - wl["flop_final"] - flop_final_x,
                        "per-graph contractions on mma.sync m16n8k8 TF32 with the 3xTF32 split (strict fp32 parity), tanh / ELU / "
                        "attention dot products on the FP32 and MUFU pipes; algorithmic FLOPs = the reference's dense GEMM count of "
                        "both networks minus the two final MLPs, at padded N (SURVEY 8d: symmetry / flag / dead-layer skipping is a "
                        "speed-up above this count, not reduced work); the kernel is issue / latency bound (profiles/README.md)",
                        mma_peak, "mma.sync TF32 (512 FMA/clk/SM measured, tools/mma_probe.cu) / 3 (3xTF32) at the observed SM clock")
        roofline_final = roof("final_edge_kernel", wl["flop_final"],
                              "tcgen05 kind::tf32, 3xTF32 split (3 MMAs per product): the pipe peak is taken as "
                              "measured bf16 / 2 (TF32 rate) / 3; algorithmic FLOPs = dense count over all N^2 pixels",
                              peak_tc / 6.0, "tcgen05 tf32 / 3 (3xTF32), from the measured bf16 peak")
        if roofline is None:
            roofline = roofline_final
        if roofline is not None:
            try:                                          # additive key; never allowed to cost the line
                roofline["mufu_ceiling"] = mufu_ceiling(wl, value, clk.get("sm_mhz") or peaks["sm_max_mhz"], world)
            except Exception as exc:
                roofline["mufu_ceiling"] = {"unavailable": repr(exc)[:120]}
    # ---- extra lines (one GPU): worst-case flags, the network evaluated without dead-chain elimination, the other BASELINE
    # configurations -- each `min(R, 3)` regions of K steps, median
    extras = None
    if world == 1 and not args.no_extra:
        extras = {}

        def quick(name, wl_, weights_, B_, mode, env=None):
            try:
                S_ = setup(wl_, weights_, B_, mode, env)
                t_, _, st_, _ = timed(S_, 10, min(R, 3))
                m_ = float(np.median(t_))
                extras[name] = {"value": rate(S_, m_), "unit": "graphs/s", "ms_per_step": m_ / K, "batch": B_, "flags": mode,
                                "mean_nodes": float(np.mean(S_["counts"])), "finite": bool(torch.isfinite(st_[3]).all()),
                                "pruned_chains": int(L.gdss_ctx_query(S_["eng"].ctx, b"pruned")),
                                "fused_path": int(L.gdss_ctx_query(S_["eng"].ctx, b"fused")), "workload": wl_["desc"]}
                del S_, st_
            except Exception as exc:
                extras[name] = {"unavailable": repr(exc)[:200]}
            torch.cuda.empty_cache()

        other_mode = "full" if args.flags == "ragged" else "ragged"
        quick(f"flags_{other_mode}", wl, weights, B, other_mode)
        quick("no_dead_chain_elimination", wl, weights, B, args.flags, {"GDSS_NO_PRUNE": "1"})
        for other in WORKLOADS:
            if other != args.workload:
                quick(other, WORKLOADS[other], load_weights(WORKLOADS[other]), WORKLOADS[other]["B"], "ragged")
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cb, cs = args.cpu_batch or wl["cpu_B"], wl["cpu_steps"]
        v, sps, cores = cpu_oracle_rate(wl, weights, cb, 1, cs, args.flags)
        cpu = {"value": v, "unit": "graphs/s", "cores": cores, "kind": "port",
               "sample": f"{cs} timed solver steps (after 1) of 1000 on B={cb} graphs, {sps:.2f} s/step, same weights/flags recipe"}
    torch_gpu = None
    if not args.no_torch_baseline and world == 1:
        try:
            del eng, S, x, adj, xm, am
            torch.cuda.empty_cache()
            tb = min(B, args.torch_batch)
            v, sps, _ = cpu_oracle_rate(wl, weights, tb, 2, 4, args.flags, device=str(dev))
            torch_gpu = {"value": v, "unit": "graphs/s", "kind": "port (oracle as PyTorch eager CUDA fp32 ops, TF32 off)",
                         "sample": f"4 timed solver steps (after 2) of 1000 on B={tb} graphs on the same B200, {sps * 1e3:.1f} ms/step",
                         "speedup_of_value": value / v}
        except Exception as exc:                                  # e.g. out of memory: reported, not fatal
            torch_gpu = {"unavailable": repr(exc)[:200]}
    path_tflops = value * 1000 * wl["evals"] * wl["flop_eval"] / 1e12
    line = {"metric": "sampled graphs/sec (1000-step reverse SDE)", "value": value, "unit": "graphs/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "tf32 (reduced-precision variant, <= 3e-2 on the raw scores)",
            "data": "synthetic flags (SURVEY 8d node-count recipe), shipped checkpoint weights",
            "config": workload_config(args, wl, B), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches) * world,
            "repeats": repeats_info, "other_scaling": other_scaling, "extra": extras,
            "pruned_chains": pruned,
            "roofline": roofline, "roofline_final_mlp": (roofline_final if not args.no_profile else None), "cpu_baseline": cpu, "torch_eager_same_gpu": torch_gpu,
            "path_algorithmic_tflops": path_tflops,
            "mean_nodes": float(np.mean(counts)), "finite": finite, "kernel_ms": kernel_ms}
    print(json.dumps(line), flush=True)
    shard.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
