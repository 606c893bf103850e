"""Host-side SDE descriptors with the reference's class names and constructor arguments
(sde.py:105-275).  On the sampling hot path an SDE is only a handful of per-step scalars; these
classes carry the parameters and the scalar algebra used to fill the coefficient table
(gdss_b200/schedule.py).  Objects of the reference's own ``sde.VPSDE / VESDE / subVPSDE`` are
accepted everywhere these are (duck-typed on ``beta_0/beta_1`` or ``sigma_min/sigma_max`` and ``N``).
"""
import numpy as np
import torch


class _SDE:
    kind = None

    def __init__(self, lo, hi, N):
        self.N = int(N)
        self._lo, self._hi = float(lo), float(hi)

    @property
    def T(self):
        return 1

    def prior_sampling(self, shape):
        """sde.py:140-141 / :201-202 -- N(0, I) on the host."""
        return torch.randn(*shape)

    def prior_sampling_sym(self, shape):
        """sde.py:143-145 / :204-207 -- strict upper triangle of a full draw, mirrored."""
        z = torch.randn(*shape).triu(1)
        return z + z.transpose(-1, -2)


class VPSDE(_SDE):
    """sde.py:105-168."""
    kind = "VP"

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(beta_min, beta_max, N)
        self.beta_0, self.beta_1 = beta_min, beta_max
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas

    def marginal_prob(self, x, t):
        """sde.py:134-138 -> (mean, std); t is a [B] tensor."""
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        shape = (-1,) + (1,) * (x.dim() - 1)
        return torch.exp(lmc).view(shape) * x, torch.sqrt(1. - torch.exp(2. * lmc))


class subVPSDE(_SDE):
    """sde.py:233-275."""
    kind = "subVP"

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(beta_min, beta_max, N)
        self.beta_0, self.beta_1 = beta_min, beta_max
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas

    def marginal_prob(self, x, t):
        """sde.py:259-263 (std = 1 - exp(2 lmc), no square root)."""
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        shape = (-1,) + (1,) * (x.dim() - 1)
        return torch.exp(lmc).view(shape) * x, 1 - torch.exp(2. * lmc)


class VESDE(_SDE):
    """sde.py:171-230."""
    kind = "VE"

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(sigma_min, sigma_max, N)
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    def marginal_prob(self, x, t):
        """sde.py:196-199."""
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t


def sde_kind(sde):
    """'VP' | 'VE' | 'subVP' for one of ours or one of the reference's SDE objects."""
    k = getattr(sde, "kind", None)
    if k in ("VP", "VE", "subVP"):
        return k
    name = type(sde).__name__
    if name in ("VPSDE", "VESDE", "subVPSDE"):
        return {"VPSDE": "VP", "VESDE": "VE", "subVPSDE": "subVP"}[name]
    raise NotImplementedError(f"SDE class {name} not yet supported.")       # utils/loader.py:107


def sde_params(sde):
    """(kind, lo, hi, N)."""
    k = sde_kind(sde)
    if k == "VE":
        return k, float(sde.sigma_min), float(sde.sigma_max), int(sde.N)
    return k, float(sde.beta_0), float(sde.beta_1), int(sde.N)


def load_sde(config_sde):
    """utils/loader.py:94-108 (config_sde has .type / .beta_min / .beta_max / .num_scales)."""
    g = (lambda k: config_sde[k]) if isinstance(config_sde, dict) else (lambda k: getattr(config_sde, k))
    t = g("type")
    if t == "VP":
        return VPSDE(beta_min=g("beta_min"), beta_max=g("beta_max"), N=g("num_scales"))
    if t == "VE":
        return VESDE(sigma_min=g("beta_min"), sigma_max=g("beta_max"), N=g("num_scales"))
    if t == "subVP":
        return subVPSDE(beta_min=g("beta_min"), beta_max=g("beta_max"), N=g("num_scales"))
    raise NotImplementedError(f"SDE class {t} not yet supported.")
