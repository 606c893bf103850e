"""Shared helpers of the GPU parity tests (test infrastructure)."""
import numpy as np
import torch

from oracle import gdss_oracle as O
from oracle.cases import NumpyDraws, case_weights, load_traj


def rel(a, b):
    """rel-L2 error; NaNs (the reference itself produces them, e.g. the last S4 half-transition of a
    VP SDE with few scales: sqrt of a negative variance) must sit at the same places."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    na, nb = torch.isnan(a), torch.isnan(b)
    if not torch.equal(na, nb):
        return float("inf")
    a, b = torch.where(na, torch.zeros_like(a), a), torch.where(nb, torch.zeros_like(b), b)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def sde_from(spec, scales):
    from gdss_b200.sde import VPSDE, VESDE, subVPSDE
    kind, lo, hi = spec
    return {"VP": VPSDE, "VE": VESDE, "subVP": subVPSDE}[kind](lo, hi, scales)


def draws_per_step(case):
    if case["predictor"] == "S4":
        return 3
    return case.get("n_steps", 1) + 1 if case["corrector"] == "Langevin" else 1


def injected_noise(case, B, N, F, iters):
    """Replays NumpyDraws(seed) in the reference's call order -> (prior_x_raw, prior_adj_raw, inj_x, inj_adj).
    Order inside a step: S4 / n_steps = 1 alternate x, adj per phase; a Langevin corrector with n_steps = K draws its K
    node-feature noises first, then its K adjacency noises (solver.py:126-147), then the predictor pair."""
    d = NumpyDraws(case["seed"])
    px, pa = d((B, N, F)), d((B, N, N))
    k = draws_per_step(case)
    K = case.get("n_steps", 1) if (case["predictor"] != "S4" and case["corrector"] == "Langevin") else 0
    ix = torch.empty(iters, k, B, N, F)
    ia = torch.empty(iters, k, B, N, N)
    for i in range(iters):
        for j in range(K):
            ix[i, j] = d((B, N, F))
        for j in range(K):
            ia[i, j] = d((B, N, N))
        for j in range(K, k):
            ix[i, j] = d((B, N, F))
            ia[i, j] = d((B, N, N))
    return px, pa, ix, ia


def case_setup(name):
    from gdss_b200.engine import Engine, make_sampler_desc
    from gdss_b200.schedule import build_table
    case, z = load_traj(name)
    sdx, px, sda, pa = case_weights(case)
    N, F, B = pa["max_node_num"], pa["max_feat_num"], case["B"]
    flags = torch.from_numpy(z["flags"])
    sx, sa = sde_from(case["sde_x"], case["scales"]), sde_from(case["sde_adj"], case["scales"])
    pf = case.get("probability_flow", False)
    table = build_table(sx, sa, case["predictor"], case["corrector"], case["snr"], case["scale_eps"], 1e-4, pf)
    sdesc = make_sampler_desc(sx, sa, case["predictor"], case["corrector"], case["snr"], case["scale_eps"], 1e-4,
                              case.get("n_steps", 1), pf, case.get("denoise", True))
    eng = Engine(sdx, sda, N, "cuda")
    return case, z, (sdx, px, sda, pa), (N, F, B), flags, table, sdesc, eng


def oracle_trace(case, weights, dims, flags, iters):
    sdx, px, sda, pa = weights
    N, F, B = dims
    sx = O.SDESpec(*case["sde_x"], case["scales"])
    sa = O.SDESpec(*case["sde_adj"], case["scales"])
    draw = NumpyDraws(case["seed"])
    trace = []
    with torch.no_grad():
        if case["predictor"] == "S4":
            O.s4_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw, snr=case["snr"],
                        scale_eps=case["scale_eps"], denoise=case.get("denoise", True), eps=1e-4, trace=trace,
                        max_iters=iters)
        else:
            O.pc_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw, predictor=case["predictor"],
                        corrector=case["corrector"], snr=case["snr"], scale_eps=case["scale_eps"],
                        n_steps=case.get("n_steps", 1), probability_flow=case.get("probability_flow", False),
                        denoise=case.get("denoise", True), eps=1e-4, trace=trace, max_iters=iters)
    return trace
