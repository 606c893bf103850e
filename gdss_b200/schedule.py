"""Per-step coefficient table of the reverse-diffusion solvers (host side, computed once).

Every batch entry of the reference carries the same time t, so all of sde.py's per-step tensor
algebra collapses to a few scalars per step.  They are evaluated here with 1-element fp32 torch
tensors in the reference's own expression order (so the scalars are the reference's fp32 values)
and folded into the affine form the update kernels apply (include/gdss_b200.h, gdss_step_coef):

    score      s      = cs * raw                                   losses.py:13-29
    Langevin   eps    = (snr * |z| / |s|)^2 * 2 * alpha            solver.py:120-134, :240-258
    predictor  y_mean = pa * y + pb * raw ;  y = y_mean + pc * z   solver.py:45-61, :72-86
    S4         y = mu1*y + sig1*z ; y += drift*raw ; y_mean = mu2*y ; y = y_mean + sig2*z   solver.py:261-277
"""
import numpy as np
import torch

from ._lib import COL, STEP_COEF_FLOATS
from .sde import sde_params


class _Scalars:
    """Scalar algebra of one SDE (kind, lo, hi, N) on 0-dim fp32 tensors."""

    def __init__(self, sde):
        self.kind, self.lo, self.hi, self.N = sde_params(sde)
        if self.kind in ("VP", "subVP"):
            self.discrete_betas = torch.linspace(self.lo / self.N, self.hi / self.N, self.N)     # sde.py:117, :245
            self.alphas = 1. - self.discrete_betas
        else:
            self.discrete_sigmas = torch.exp(torch.linspace(np.log(self.lo), np.log(self.hi), self.N))   # sde.py:182

    def k(self, t):
        return int((t * (self.N - 1) / 1).long())                       # sde.py:155, :216 (T = 1)

    def beta_t(self, t):
        return self.lo + t * (self.hi - self.lo)                        # sde.py:128, :253

    def diffusion(self, t):
        if self.kind == "VP":
            return torch.sqrt(self.beta_t(t))                           # sde.py:130
        if self.kind == "subVP":
            discount = 1. - torch.exp(-2 * self.lo * t - (self.hi - self.lo) * t ** 2)
            return torch.sqrt(self.beta_t(t) * discount)                # sde.py:255-256
        sigma = self.lo * (self.hi / self.lo) ** t
        return sigma * torch.sqrt(torch.tensor(2 * (np.log(self.hi) - np.log(self.lo))))    # sde.py:190-193

    def drift_coeff(self, t):
        return torch.zeros(()) if self.kind == "VE" else -0.5 * self.beta_t(t)              # sde.py:129, :191, :254

    def score_scale(self, t):
        if self.kind == "VE":
            return torch.ones(())                                       # losses.py:24-29
        lmc = -0.25 * t ** 2 * (self.hi - self.lo) - 0.5 * t * self.lo  # sde.py:135, :260
        std = torch.sqrt(1. - torch.exp(2. * lmc)) if self.kind == "VP" else 1 - torch.exp(2. * lmc)
        return -1.0 / std                                               # losses.py:16-20

    def langevin_alpha(self, t, s4):
        # solver.py:120-124 (VPSDE or subVPSDE) vs solver.py:240-243 (S4: isinstance(sde, VPSDE) only)
        if self.kind == "VP" or (self.kind == "subVP" and not s4):
            return self.alphas[self.k(t)]
        return torch.ones(())

    def reverse_step(self, t):
        """(f_coeff, G) with f = f_coeff * y  (sde.py:153-161, :214-222, :48-62)."""
        if self.kind == "VP":
            kk = self.k(t)
            return torch.sqrt(self.alphas[kk]) - 1.0, torch.sqrt(self.discrete_betas[kk])
        if self.kind == "VE":
            kk = self.k(t)
            adjacent = torch.zeros(()) if kk == 0 else self.discrete_sigmas[kk - 1]
            return torch.zeros(()), torch.sqrt(self.discrete_sigmas[kk] ** 2 - adjacent ** 2)
        dt = 1 / self.N
        return self.drift_coeff(t) * dt, self.diffusion(t) * torch.sqrt(torch.tensor(dt))

    def transition(self, t, dt):
        """(mean_coeff, std) of sde.py:163-168 / :224-230."""
        if self.kind == "VP":
            lmc = 0.25 * dt * (2 * self.lo + (2 * t + dt) * (self.hi - self.lo))
            return torch.exp(-lmc), torch.sqrt(1. - torch.exp(2. * lmc))
        if self.kind == "VE":
            std = torch.square(self.lo * (self.hi / self.lo) ** t) - torch.square(self.lo * (self.hi / self.lo) ** (t + dt))
            return torch.ones(()), torch.sqrt(std)
        raise AttributeError("'subVPSDE' object has no attribute 'transition'")     # as the reference behaves


def build_table(sde_x, sde_adj, predictor, corrector, snr, scale_eps, eps, probability_flow=False):
    """-> float32 numpy array [sde_adj.N, 24] laid out as gdss_step_coef rows.

    predictor in {'Euler', 'Reverse', 'S4'}; anything else maps to Euler, any corrector other than
    'Langevin' to none (solver.py:163-164)."""
    pf = 0.5 if (probability_flow and predictor != "S4") else 1.0        # sde.py:89-91, :98-99 (S4 never builds an RSDE)
    sx, sa = _Scalars(sde_x), _Scalars(sde_adj)
    steps = sa.N
    ts = torch.linspace(1, eps, steps)                                   # solver.py:180, :218
    tab = np.zeros((steps, STEP_COEF_FLOATS), np.float32)
    s4 = predictor == "S4"
    langevin = s4 or corrector == "Langevin"
    for i in range(steps):
        t = ts[i]
        row = tab[i]
        row[COL["snr"]], row[COL["seps"]], row[COL["t"]] = snr, scale_eps, float(t)
        for d, sd in enumerate((sx, sa)):
            cs = sd.score_scale(t)
            row[COL["cs"] + d] = float(cs)
            row[COL["alpha"] + d] = 1.0
            if langevin:
                if s4:      # solver.py:240-243, :252-255: the index always comes from sde_x.N
                    kk = int((t * (sx.N - 1) / 1).long())
                    row[COL["alpha"] + d] = float(sd.alphas[kk]) if sd.kind == "VP" else 1.0
                else:
                    row[COL["alpha"] + d] = float(sd.langevin_alpha(t, False))
            if s4:
                dt = -1. / steps                                         # solver.py:220
                hdt = torch.ones(()) * (dt / 2)
                g = sd.diffusion(t)
                row[COL["drift"] + d] = float(-g ** 2 * cs * dt)         # solver.py:231-232, :268-269
                m1, s1 = sd.transition(t, hdt)
                m2, s2 = sd.transition(t + hdt, hdt)
                row[COL["mu1"] + d], row[COL["sig1"] + d] = float(m1), float(s1)
                row[COL["mu2"] + d], row[COL["sig2"] + d] = float(m2), float(s2)
            elif predictor == "Reverse":
                fc, G = sd.reverse_step(t)                               # solver.py:72-86, sde.py:94-100
                row[COL["pa"] + d] = float(1.0 - fc)
                row[COL["pb"] + d] = float(G ** 2 * cs * pf)
                row[COL["pc"] + d] = float(G) if pf == 1.0 else 0.0      # rev_G = 0 for the probability-flow ODE
            else:
                dt = -1. / sd.N                                          # solver.py:45-61, sde.py:85-92
                g = sd.diffusion(t)
                row[COL["pa"] + d] = float(1.0 + sd.drift_coeff(t) * dt)
                row[COL["pb"] + d] = float(-(g ** 2) * cs * dt * pf)
                row[COL["pc"] + d] = float(g * np.sqrt(-dt)) if pf == 1.0 else 0.0
    return tab
