"""Shared definitions for the golden fixtures (test infrastructure only).

The case tables are the single source of truth for oracle/make_golden.py (which
runs the real reference) and for the tests that replay the fixtures through the
oracle and through the CUDA path.
"""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# fixture name -> path below <reference>/checkpoints (without .pth)
CKPTS = {
    "qm9": "QM9/gdss_qm9",
    "zinc250k": "ZINC250k/gdss_zinc250k",
    "zinc250k_v2": "ZINC250k/gdss_zinc250k_v2",
    "community_small": "community_small/gdss_community_small",
    "ego_small": "ego_small/gdss_ego_small",
    "enzymes": "ENZYMES/gdss_enzymes",
    "grid": "grid/gdss_grid",
}


def _case(name, ckpt, sde_x, sde_adj, predictor, corrector, snr, seps, scales, B, seed, **kw):
    d = dict(name=name, ckpt=ckpt, sde_x=list(sde_x), sde_adj=list(sde_adj), predictor=predictor,
             corrector=corrector, snr=snr, scale_eps=seps, scales=scales, B=B, seed=seed)
    d.update(kw)
    return d


# Reduced number of scales (the sampler always runs sde_adj.N steps, solver.py:179): cheap runs that
# still exercise every predictor / corrector / SDE combination the shipped configs use.
SHORT_CASES = [
    _case("qm9_s4_n30", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "S4", "None", .2, .7, 30, 4, 101),
    _case("qm9_rev_lang_n30", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "Reverse", "Langevin", .2, .7, 30, 4, 102),
    _case("zinc_v2_rev_lang_n12", "zinc250k_v2", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "Langevin", .2, .8, 12, 3, 103),
    _case("zinc_v1_rev_lang_n12", "zinc250k", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "Langevin", .2, .9, 12, 3, 104),
    _case("community_euler_lang_n25", "community_small", ("VP", .1, 1.), ("VP", .1, 1.), "Euler", "Langevin", .05, .7, 25, 4, 105),
    _case("ego_euler_none_n25", "ego_small", ("VP", .1, 1.), ("VP", .1, 1.), "Euler", "None", .0, .0, 25, 4, 106),
    _case("community_subvp_rev_lang_n10", "community_small", ("subVP", .1, 1.), ("VP", .1, 1.), "Reverse", "Langevin", .05, .7, 10, 3, 107),
    _case("community_rev_none_nodenoise_n10", "community_small", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "None", .05, .7, 10, 3, 108, denoise=False),
    _case("enzymes_s4_n6", "enzymes", ("VP", .1, 1.), ("VE", .2, 1.), "S4", "None", .15, .7, 6, 3, 109, ema=True),
    _case("grid_rev_lang_n3", "grid", ("VP", .1, 1.), ("VP", .2, .8), "Reverse", "Langevin", .1, .7, 3, 1, 110, ema=True),
    # probability-flow ODE (sde.py:98-99): reverse drift halved, no predictor noise (the draws are still consumed).  Only the
    # Reverse predictor exists with it: EulerMaruyamaPredictor + probability_flow raises in the reference (solver.py:52
    # indexes the float 0. that sde.py:91 returns as the diffusion)
    _case("community_rev_none_pf_n10", "community_small", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "None", .0, .0, 10, 4, 111,
          probability_flow=True),
    _case("qm9_rev_lang_pf_n10", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "Reverse", "Langevin", .2, .7, 10, 4, 112,
          probability_flow=True),
    # Langevin corrector with n_steps = 2 inner iterations (solver.py:126-147): the x corrector iterates on (x_k, adj), then
    # the adjacency corrector on (x_0, adj_k)
    _case("qm9_rev_lang_ns2_n10", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "Reverse", "Langevin", .2, .7, 10, 4, 113, n_steps=2),
    _case("community_euler_lang_ns3_n8", "community_small", ("VP", .1, 1.), ("VP", .1, 1.), "Euler", "Langevin", .05, .7, 8, 3,
          114, n_steps=3),
]

# Full 1000-scale runs (the BASELINE.json configurations at a small batch).
FULL_CASES = [
    _case("qm9_s4_full", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "S4", "None", .2, .7, 1000, 4, 201),
    _case("qm9_rev_lang_full", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "Reverse", "Langevin", .2, .7, 1000, 4, 202),
    _case("community_euler_lang_full", "community_small", ("VP", .1, 1.), ("VP", .1, 1.), "Euler", "Langevin", .05, .7, 1000, 3, 203),
    _case("zinc_v2_rev_lang_full", "zinc250k_v2", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "Langevin", .2, .8, 1000, 2, 204),
    # larger batches for the bit-exact thresholded-output claim (32 molecules each) and a full-length C4 run
    _case("qm9_s4_full_b32", "qm9", ("VE", .1, 1.), ("VE", .1, 1.), "S4", "None", .2, .7, 1000, 32, 205),
    _case("zinc_v2_rev_lang_full_b32", "zinc250k_v2", ("VP", .1, 1.), ("VE", .2, 1.), "Reverse", "Langevin", .2, .8, 1000, 32, 206),
    _case("enzymes_s4_full_b4", "enzymes", ("VP", .1, 1.), ("VE", .2, 1.), "S4", "None", .15, .7, 1000, 4, 207, ema=True),
]

ALL_CASES = {c["name"]: c for c in SHORT_CASES + FULL_CASES}


class NumpyDraws:
    """Raw N(0,1) draws from numpy's frozen legacy stream ``RandomState(seed)``
    (bit-stable across numpy versions), fp32.  ``tape`` keeps every draw."""

    def __init__(self, seed: int, keep: bool = False):
        self.rs = np.random.RandomState(seed)
        self.keep = keep
        self.tape = []

    def __call__(self, shape):
        t = torch.from_numpy(self.rs.standard_normal(tuple(shape)).astype(np.float32))
        if self.keep:
            self.tape.append(t)
        return t


def plain(obj):
    """EasyDict / dict tree -> plain python containers."""
    if isinstance(obj, dict):
        return {str(k): plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [plain(v) for v in obj]
    return obj


def load_golden_ckpt(name: str) -> dict:
    return torch.load(os.path.join(GOLD, "ckpt", f"{name}.pt"), map_location="cpu", weights_only=True)


def load_kat(name: str):
    return np.load(os.path.join(GOLD, f"kat_{name}.npz"))


def load_traj(name: str):
    z = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
    return json.loads(str(z["case"])), z


def case_weights(case):
    """(sd_x, p_x, sd_a, p_a) for a case, EMA applied where the config asks (oracle-side)."""
    from oracle import gdss_oracle as O
    ck = load_golden_ckpt(case["ckpt"])
    sdx, sda = O.strip_module_prefix(ck["x_state_dict"]), O.strip_module_prefix(ck["adj_state_dict"])
    if case.get("ema"):
        sdx, sda = O.apply_ema(sdx, ck["ema_x"]), O.apply_ema(sda, ck["ema_adj"])
    return sdx, ck["params_x"], sda, ck["params_adj"]
