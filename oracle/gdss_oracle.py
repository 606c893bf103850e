"""CPU oracle for the GDSS reverse-diffusion sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``gdss_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker / the timed CPU arm.

This is a from-scratch *functional restatement* (plain torch fp32 on the CPU,
driven by raw ``state_dict`` tensors, no ``nn.Module``) of the algorithm the
reference implements in

  * ``solver.py``      predictors / correctors / ``get_pc_sampler`` / ``S4_solver``
  * ``sde.py``         VPSDE / VESDE / subVPSDE scalar algebra
  * ``losses.py``      ``get_score_fn`` score scaling
  * ``models/*.py``    ScoreNetworkX / ScoreNetworkX_GMH / ScoreNetworkA
  * ``utils/graph_utils.py``  masks, ``gen_noise``, ``pow_tensor``, ``quantize*``

Every function cites the reference ``file:line`` it follows (paths relative to
the upstream repository root).

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is therefore pinned against *outputs of the
unmodified reference itself*, generated in the build container by
``oracle/make_golden.py`` (which imports ``/root/reference``) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

Randomness: the oracle never draws noise itself.  Every Gaussian draw is
requested from a ``draw(shape) -> tensor`` callable in exactly the order the
reference calls ``torch.randn``/``torch.randn_like`` so that (a) seeding the
global torch CPU generator reproduces the reference bit-for-bit and (b) the
CUDA path can be fed the very same raw draws ("injected noise").
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]
Draw = Callable[[Sequence[int]], Tensor]


# --------------------------------------------------------------------------
# tensor helpers  (utils/graph_utils.py)
# --------------------------------------------------------------------------
def mask_x(x: Tensor, flags: Optional[Tensor]) -> Tensor:
    """utils/graph_utils.py:8-12 -- zero the feature rows of padded nodes."""
    if flags is None:
        return x
    return x * flags.unsqueeze(-1)


def mask_adjs(a: Tensor, flags: Optional[Tensor]) -> Tensor:
    """utils/graph_utils.py:16-29 -- zero rows then columns of padded nodes
    (two successive multiplies, [B,N,N] or [B,C,N,N])."""
    if flags is None:
        return a
    f = flags if a.dim() == 3 else flags.unsqueeze(1)
    a = a * f.unsqueeze(-1)
    a = a * f.unsqueeze(-2)
    return a


def node_flags(adj: Tensor, eps: float = 1e-5) -> Tensor:
    """utils/graph_utils.py:33-39."""
    fl = (adj.abs().sum(-1) > eps).to(torch.float32)
    if fl.dim() == 3:
        fl = fl[:, 0, :]
    return fl


def sym_noise_from_raw(raw: Tensor, flags: Tensor) -> Tensor:
    """utils/graph_utils.py:80-83 -- strict upper triangle of a full N x N draw,
    mirrored, then masked."""
    z = raw.triu(1)
    z = z + z.transpose(-1, -2)
    return mask_adjs(z, flags)


def gen_noise(like: Tensor, flags: Tensor, sym: bool, draw: Draw) -> Tensor:
    """utils/graph_utils.py:78-86 (the draw is ``randn_like`` = full shape)."""
    raw = draw(tuple(like.shape))
    return sym_noise_from_raw(raw, flags) if sym else mask_x(raw, flags)


def pow_stack(adj: Tensor, cnum: int) -> Tensor:
    """utils/graph_utils.py:133-142 -- [A, A^2, ..., A^cnum] on a channel axis;
    each power is (previous power) @ A."""
    chans = [adj]
    cur = adj
    for _ in range(cnum - 1):
        cur = torch.bmm(cur, adj)
        chans.append(cur)
    return torch.stack(chans, dim=1)


def quantize(adjs: Tensor, thr: float = 0.5) -> Tensor:
    """utils/graph_utils.py:90-92."""
    return (adjs >= thr).to(adjs.dtype)


def quantize_mol(adjs: Tensor) -> np.ndarray:
    """utils/graph_utils.py:97-106 -- bond-order bins {0,1,2,3} with edges at
    0.5 / 1.5 / 2.5 (>= lower edge, < upper edge); int64 numpy out."""
    a = adjs.detach().cpu()
    out = torch.zeros_like(a, dtype=torch.int64)
    out[a >= 0.5] = 1
    out[a >= 1.5] = 2
    out[a >= 2.5] = 3
    return out.numpy()


def decode_mol(x: Tensor, adj: Tensor) -> Tuple[Tensor, Tensor]:
    """sampler.py:131-138 -- what Sampler_mol does with the sampler's output before gen_mol: quantize_mol, bond classes
    remapped (0 no bond, 1 S, 2 D, 3 T) -> (3, 0, 1, 2), one-hot over 4 classes as [B,4,N,N]; atoms x > 0.5 with the
    virtual-atom channel 1 - sum appended.  -> (x [B,N,F+1] int64, adj [B,4,N,N] int64)."""
    q = torch.from_numpy(quantize_mol(adj)) - 1
    q[q == -1] = 3
    a = torch.nn.functional.one_hot(q, num_classes=4).permute(0, 3, 1, 2)
    xi = torch.where(x > 0.5, 1, 0)
    xi = torch.concat([xi, 1 - xi.sum(dim=-1, keepdim=True)], dim=-1)
    return xi, a


# --------------------------------------------------------------------------
# layers  (models/layers.py, models/attention.py)
# --------------------------------------------------------------------------
def dense_gcn(x: Tensor, adj: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """models/layers.py:49-88 (add_loop=True, improved=False, mask=None).

    diag := 1 (overwrite), deg = rowsum clamped to >= 1, symmetric
    normalisation, out = A_hat @ (x @ W) + b.  ``weight`` is [in, out]."""
    n = adj.shape[-1]
    a = adj.clone()
    ii = torch.arange(n, device=adj.device)
    a[:, ii, ii] = 1.0
    xw = torch.matmul(x, weight)
    dis = a.sum(dim=-1).clamp(min=1).pow(-0.5)
    a = dis.unsqueeze(-1) * a * dis.unsqueeze(-2)
    return torch.matmul(a, xw) + bias


def elu(h: Tensor) -> Tensor:
    return torch.nn.functional.elu(h)


def mlp(sd: StateDict, prefix: str, h: Tensor, num_layers: int, act=elu) -> Tensor:
    """models/layers.py:135-153 with use_bn=False: Linear, act, ..., Linear.
    ``nn.Linear`` weights are [out, in]."""
    if num_layers == 1:
        return torch.nn.functional.linear(h, sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])
    for i in range(num_layers - 1):
        h = act(torch.nn.functional.linear(h, sd[f"{prefix}.linears.{i}.weight"],
                                           sd[f"{prefix}.linears.{i}.bias"]))
    i = num_layers - 1
    return torch.nn.functional.linear(h, sd[f"{prefix}.linears.{i}.weight"], sd[f"{prefix}.linears.{i}.bias"])


def attention(sd: StateDict, prefix: str, x: Tensor, adj_c: Tensor, num_heads: int) -> Tuple[Tensor, Tensor]:
    """models/attention.py:25-51 with conv='GCN', attention_mask=None.

    Returns (V [B,N,out], M [B,N,N]); M = sym(mean_h tanh(Q_h K_h^T / sqrt(out_dim)))
    -- scaled by sqrt(out_dim) (= V width), not by the head width."""
    q = dense_gcn(x, adj_c, sd[f"{prefix}.gnn_q.weight"], sd[f"{prefix}.gnn_q.bias"])
    k = dense_gcn(x, adj_c, sd[f"{prefix}.gnn_k.weight"], sd[f"{prefix}.gnn_k.bias"])
    v = dense_gcn(x, adj_c, sd[f"{prefix}.gnn_v.weight"], sd[f"{prefix}.gnn_v.bias"])
    b, n, attn_dim = q.shape
    out_dim = v.shape[-1]
    hd = attn_dim // num_heads
    qh = q.view(b, n, num_heads, hd).permute(2, 0, 1, 3)     # [H,B,N,hd]
    kh = k.view(b, n, num_heads, hd).permute(2, 0, 1, 3)
    s = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(out_dim)
    m = torch.tanh(s).mean(dim=0)
    m = (m + m.transpose(-1, -2)) / 2
    return v, m


def attention_layer(sd: StateDict, prefix: str, x: Tensor, adjc: Tensor, flags: Tensor,
                    num_linears: int, num_heads: int) -> Tuple[Tensor, Tensor]:
    """models/attention.py:93-116.

    x [B,N,Fi], adjc [B,Ci,N,N] -> (x_out [B,N,nhid], adj_out [B,Co,N,N])."""
    c_in = adjc.shape[1]
    vs, ms = [], []
    for c in range(c_in):
        v, m = attention(sd, f"{prefix}.attn.{c}", x, adjc[:, c], num_heads)
        vs.append(v)
        ms.append(m)
    x_out = torch.tanh(mask_x(mlp(sd, f"{prefix}.multi_channel", torch.cat(vs, dim=-1), 2), flags))
    edge_in = torch.cat([torch.stack(ms, dim=-1), adjc.permute(0, 2, 3, 1)], dim=-1)   # [B,N,N,2Ci]
    bsz, n = edge_in.shape[0], edge_in.shape[1]
    s = mlp(sd, f"{prefix}.mlp", edge_in.reshape(-1, 2 * c_in), num_linears)             # rows = (b,i,j)
    s = s.view(bsz, n, n, -1).permute(0, 3, 1, 2)                                       # [B,Co,N,N]
    s = s + s.transpose(-1, -2)
    return x_out, mask_adjs(s, flags)


# --------------------------------------------------------------------------
# score networks  (models/ScoreNetwork_X.py, models/ScoreNetwork_A.py)
# --------------------------------------------------------------------------
def score_network_x(sd: StateDict, p: dict, x: Tensor, adj: Tensor, flags: Tensor) -> Tensor:
    """models/ScoreNetwork_X.py:32-46 -- depth x tanh(GCN) on the raw adjacency,
    concat [x, x1..xd], 3-layer ELU MLP, mask."""
    feats = [x]
    h = x
    for l in range(p["depth"]):
        h = torch.tanh(dense_gcn(h, adj, sd[f"layers.{l}.weight"], sd[f"layers.{l}.bias"]))
        feats.append(h)
    out = mlp(sd, "final", torch.cat(feats, dim=-1), 3)
    return mask_x(out, flags)


def score_network_x_gmh(sd: StateDict, p: dict, x: Tensor, adj: Tensor, flags: Tensor) -> Tensor:
    """models/ScoreNetwork_X.py:75-89 -- AttentionLayer stack on the power stack,
    an extra tanh on every layer's (already tanh'd) node output."""
    adjc = pow_stack(adj, p["c_init"])
    feats = [x]
    h = x
    for l in range(p["depth"]):
        h, adjc = attention_layer(sd, f"layers.{l}", h, adjc, flags, p["num_linears"], p.get("num_heads", 4))
        h = torch.tanh(h)
        feats.append(h)
    out = mlp(sd, "final", torch.cat(feats, dim=-1), 3)
    return mask_x(out, flags)


def score_network_a(sd: StateDict, p: dict, x: Tensor, adj: Tensor, flags: Tensor) -> Tensor:
    """models/ScoreNetwork_A.py:131-150 -- power stack, L attention layers, concat
    of every channel stack, per-edge 3-layer ELU MLP -> 1, zero diagonal, mask."""
    adjc = pow_stack(adj, p["c_init"])
    stacks = [adjc]
    h = x
    for l in range(p["num_layers"]):
        h, adjc = attention_layer(sd, f"layers.{l}", h, adjc, flags, p["num_linears"], p.get("num_heads", 4))
        stacks.append(adjc)
    cat = torch.cat(stacks, dim=1).permute(0, 2, 3, 1)         # [B,N,N,fdim]
    score = mlp(sd, "final", cat, 3).squeeze(-1)
    n = adj.shape[-1]
    score = score * (1.0 - torch.eye(n, device=score.device)).unsqueeze(0)
    return mask_adjs(score, flags)


def run_model(sd: StateDict, params: dict, x: Tensor, adj: Tensor, flags: Tensor) -> Tensor:
    """utils/loader.py:37-48 dispatch on ``model_type``."""
    kind = params["model_type"]
    if kind == "ScoreNetworkX":
        return score_network_x(sd, params, x, adj, flags)
    if kind == "ScoreNetworkX_GMH":
        return score_network_x_gmh(sd, params, x, adj, flags)
    if kind == "ScoreNetworkA":
        return score_network_a(sd, params, x, adj, flags)
    raise ValueError(f"Model Name <{kind}> is Unknown")


def strip_module_prefix(sd: StateDict) -> StateDict:
    """utils/loader.py:188-190."""
    first = next(iter(sd.keys()))
    if "module." in first:
        return {k[7:]: v for k, v in sd.items()}
    return dict(sd)


def apply_ema(sd: StateDict, ema_state: dict) -> StateDict:
    """sampler.py:47-52 + utils/ema.py:37-47 -- shadow params replace the
    parameters in ``parameters()`` order (== state_dict order here: the score
    networks register no buffers)."""
    keys = list(sd.keys())
    shadow = ema_state["shadow_params"]
    assert len(keys) == len(shadow)
    return {k: s.detach().clone() for k, s in zip(keys, shadow)}


# --------------------------------------------------------------------------
# SDE scalar algebra  (sde.py).  Every batch entry carries the same t, so all
# of this is evaluated on 1-element fp32 tensors with the reference's own
# torch expressions and returned as python floats (exact fp32 values).
# --------------------------------------------------------------------------
class SDESpec:
    """kind in {'VP','VE','subVP'}; (lo, hi) = (beta_min, beta_max) or
    (sigma_min, sigma_max) exactly as utils/loader.py:94-108 passes them."""

    def __init__(self, kind: str, lo: float, hi: float, N: int = 1000):
        if kind not in ("VP", "VE", "subVP"):
            raise NotImplementedError(f"SDE class {kind} not yet supported.")
        self.kind, self.lo, self.hi, self.N = kind, lo, hi, N
        self.T = 1
        if kind in ("VP", "subVP"):
            self.discrete_betas = torch.linspace(lo / N, hi / N, N)       # sde.py:117 / :245
            self.alphas = 1.0 - self.discrete_betas                       # sde.py:118 / :246
        else:
            self.discrete_sigmas = torch.exp(torch.linspace(np.log(lo), np.log(hi), N))   # sde.py:182

    # t is a 0-dim / 1-element fp32 tensor below ---------------------------
    def timestep(self, t: Tensor) -> int:
        return int((t * (self.N - 1) / self.T).long())                    # sde.py:155 / :216

    def diffusion(self, t: Tensor) -> Tensor:
        """second return value of ``sde(x, t)``: sde.py:127-131, :189-194, :252-257."""
        if self.kind == "VP":
            return torch.sqrt(self.lo + t * (self.hi - self.lo))
        if self.kind == "subVP":
            beta_t = self.lo + t * (self.hi - self.lo)
            discount = 1.0 - torch.exp(-2 * self.lo * t - (self.hi - self.lo) * t ** 2)
            return torch.sqrt(beta_t * discount)
        sigma = self.lo * (self.hi / self.lo) ** t
        return sigma * torch.sqrt(torch.tensor(2 * (np.log(self.hi) - np.log(self.lo))))

    def drift_coeff(self, t: Tensor) -> Tensor:
        """drift = coeff * x : -0.5*beta(t) for VP/subVP, 0 for VE."""
        if self.kind == "VE":
            return torch.zeros_like(t)
        return -0.5 * (self.lo + t * (self.hi - self.lo))

    def marginal_std(self, t: Tensor) -> Tensor:
        """sde.py:134-138, :196-199, :259-263 (subVP returns 1-exp(..), no sqrt)."""
        if self.kind == "VE":
            return self.lo * (self.hi / self.lo) ** t
        lmc = -0.25 * t ** 2 * (self.hi - self.lo) - 0.5 * t * self.lo
        if self.kind == "VP":
            return torch.sqrt(1.0 - torch.exp(2.0 * lmc))
        return 1 - torch.exp(2.0 * lmc)

    def discretize(self, y: Tensor, t: Tensor) -> Tuple[Tensor, Tensor]:
        """(f, G).  VP: sde.py:153-161 (f = sqrt(alpha)*y - y); VE: :214-222
        (f = 0, sigma_{-1} := 0); subVP inherits the Euler rule sde.py:48-62."""
        if self.kind == "VP":
            k = self.timestep(t)
            beta = self.discrete_betas[k]
            alpha = self.alphas[k]
            return torch.sqrt(alpha) * y - y, torch.sqrt(beta)
        if self.kind == "VE":
            k = self.timestep(t)
            sigma = self.discrete_sigmas[k]
            adjacent = torch.zeros(()) if k == 0 else self.discrete_sigmas[k - 1]
            return torch.zeros_like(y), torch.sqrt(sigma ** 2 - adjacent ** 2)
        dt = 1 / self.N
        return (self.drift_coeff(t) * y) * dt, self.diffusion(t) * torch.sqrt(torch.tensor(dt))

    def transition(self, t: Tensor, dt: Tensor) -> Tuple[Tensor, Tensor]:
        """(mean_coeff, std), negative dt.  VP: sde.py:163-168; VE: :224-230."""
        if self.kind == "VP":
            lmc = 0.25 * dt * (2 * self.lo + (2 * t + dt) * (self.hi - self.lo))
            return torch.exp(-lmc), torch.sqrt(1.0 - torch.exp(2.0 * lmc))
        if self.kind == "VE":
            std = torch.square(self.lo * (self.hi / self.lo) ** t) - \
                torch.square(self.lo * (self.hi / self.lo) ** (t + dt))
            return torch.ones(()), torch.sqrt(std)
        raise AttributeError("subVPSDE has no transition")               # as in the reference

    def langevin_alpha(self, t: Tensor, s4: bool) -> Tensor:
        """solver.py:120-124 (VP and subVP) vs solver.py:240-243 (S4: VP only)."""
        if self.kind == "VP" or (self.kind == "subVP" and not s4):
            return self.alphas[self.timestep(t)]
        return torch.ones(())


def make_score_fn(sde: SDESpec, sd: StateDict, params: dict):
    """losses.py:6-34 with train=False, continuous=True."""
    if sde.kind in ("VP", "subVP"):
        def fn(x, adj, flags, t):
            return -run_model(sd, params, x, adj, flags) / sde.marginal_std(t)
    else:
        def fn(x, adj, flags, t):
            return run_model(sd, params, x, adj, flags)
    return fn


# --------------------------------------------------------------------------
# solver  (solver.py)
# --------------------------------------------------------------------------
def dsm_loss(sde_x: "SDESpec", sde_adj: "SDESpec", sd_x: StateDict, p_x: dict, sd_a: StateDict, p_a: dict, x: Tensor,
             adj: Tensor, t: Tensor, raw_x: Tensor, raw_adj: Tensor, reduce_mean: bool = False) -> Tuple[Tensor, Tensor]:
    """losses.py:37-81 (likelihood_weighting = False) with the three random draws supplied: t [B] (losses.py:47),
    raw_x / raw_adj = the randn_like draws inside gen_noise (losses.py:50, :55)."""
    def marginal(sde, y, tt):                       # sde.py:134-138 / :196-199 / :259-263 with a [B] time vector
        if sde.kind == "VE":
            return y, sde.lo * (sde.hi / sde.lo) ** tt
        lmc = -0.25 * tt ** 2 * (sde.hi - sde.lo) - 0.5 * tt * sde.lo
        mean = torch.exp(lmc).view((-1,) + (1,) * (y.dim() - 1)) * y
        return mean, (torch.sqrt(1. - torch.exp(2. * lmc)) if sde.kind == "VP" else 1 - torch.exp(2. * lmc))

    def score(sde, sd, p, px, pa, fl):              # losses.py:6-34
        net = run_model(sd, p, px, pa, fl)
        if sde.kind == "VE":
            return net
        std = marginal(sde, torch.zeros_like(pa), t)[1]
        return -net / std.view((-1,) + (1,) * (net.dim() - 1))

    reduce_op = torch.mean if reduce_mean else (lambda v, dim: 0.5 * torch.sum(v, dim=dim))
    flags = node_flags(adj)
    z_x = mask_x(raw_x, flags)
    mean_x, std_x = marginal(sde_x, x, t)
    pert_x = mask_x(mean_x + std_x[:, None, None] * z_x, flags)
    z_adj = sym_noise_from_raw(raw_adj, flags)
    mean_a, std_a = marginal(sde_adj, adj, t)
    pert_a = mask_adjs(mean_a + std_a[:, None, None] * z_adj, flags)
    s_x = score(sde_x, sd_x, p_x, pert_x, pert_a, flags)
    s_a = score(sde_adj, sd_a, p_a, pert_x, pert_a, flags)
    l_x = torch.square(s_x * std_x[:, None, None] + z_x)
    l_x = reduce_op(l_x.reshape(l_x.shape[0], -1), dim=-1)
    l_a = torch.square(s_a * std_a[:, None, None] + z_adj)
    l_a = reduce_op(l_a.reshape(l_a.shape[0], -1), dim=-1)
    return torch.mean(l_x), torch.mean(l_a)


def _batch_mean_norm(v: Tensor) -> Tensor:
    """solver.py:130-131 -- per-graph Frobenius norm, then the mean over the batch."""
    return torch.norm(v.reshape(v.shape[0], -1), dim=-1).mean()


def langevin_update(state: Tensor, score: Tensor, noise: Tensor, snr: float, seps: float,
                    alpha: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """solver.py:130-134 / :141-145 / :238-247 / :250-258.  Returns (new, mean, step_size)."""
    gn = _batch_mean_norm(score)
    nn_ = _batch_mean_norm(noise)
    step = (snr * nn_ / gn) ** 2 * 2 * alpha
    mean = state + step * score
    new = mean + torch.sqrt(step * 2) * noise * seps
    return new, mean, step


def prior(shape_x, shape_adj, flags: Tensor, draw: Draw) -> Tuple[Tensor, Tensor]:
    """solver.py:174-178 -- x draw first, then the adjacency draw; triu(1)+T; masks."""
    x = draw(tuple(shape_x))
    raw = draw(tuple(shape_adj))
    a = raw.triu(1)
    a = a + a.transpose(-1, -2)
    return mask_x(x, flags), mask_adjs(a, flags)


def pc_sample(sde_x: SDESpec, sde_adj: SDESpec, sd_x: StateDict, p_x: dict, sd_a: StateDict, p_a: dict,
              shape_x, shape_adj, flags: Tensor, draw: Draw, predictor: str = "Euler",
              corrector: str = "None", snr: float = 0.1, scale_eps: float = 1.0, n_steps: int = 1,
              probability_flow: bool = False, denoise: bool = True, eps: float = 1e-3,
              trace: Optional[List[dict]] = None, max_iters: Optional[int] = None):
    """solver.py:153-196.  ``trace`` (if a list) receives the per-phase tensors."""
    sfx = make_score_fn(sde_x, sd_x, p_x)
    sfa = make_score_fn(sde_adj, sd_a, p_a)
    use_reverse = predictor == "Reverse"          # anything else -> Euler (solver.py:163)
    use_langevin = corrector == "Langevin"        # anything else -> None  (solver.py:164)
    pf = 0.5 if probability_flow else 1.0

    x, adj = prior(shape_x, shape_adj, flags, draw)
    steps = sde_adj.N
    ts = torch.linspace(sde_adj.T, eps, steps)
    x_mean, adj_mean = x, adj
    for i in range(steps if max_iters is None else min(steps, max_iters)):
        t = ts[i]
        rec = {} if trace is not None else None
        # ---- corrector on (x, adj): both updates see the pre-step x  (solver.py:187-189)
        if use_langevin:
            x0 = x
            for _ in range(n_steps):
                g = sfx(x, adj, flags, t)
                z = gen_noise(x, flags, False, draw)
                x, x_mean, ex = langevin_update(x, g, z, snr, scale_eps, sde_x.langevin_alpha(t, False))
                if rec is not None:
                    rec.update(c_score_x=g, c_eps_x=ex)
            for _ in range(n_steps):
                g = sfa(x0, adj, flags, t)
                z = gen_noise(adj, flags, True, draw)
                adj, adj_mean, ea = langevin_update(adj, g, z, snr, scale_eps, sde_adj.langevin_alpha(t, False))
                if rec is not None:
                    rec.update(c_score_adj=g, c_eps_adj=ea)
            if rec is not None:
                rec.update(c_x=x, c_adj=adj)
        # ---- predictor on (x', adj'): adj update sees the post-corrector x  (solver.py:191-193)
        x1 = x
        if use_reverse:
            # solver.py:72-86 + sde.py:94-100 : score first, then the noise draw
            f, G = sde_x.discretize(x1, t)
            s = sfx(x1, adj, flags, t)
            rev_f = f - G ** 2 * s * pf
            z = gen_noise(x1, flags, False, draw)
            x_mean = x1 - rev_f
            x = x_mean + (torch.zeros(()) if probability_flow else G) * z
            if rec is not None:
                rec.update(p_score_x=s)
            f, G = sde_adj.discretize(adj, t)
            s = sfa(x1, adj, flags, t)
            rev_f = f - G ** 2 * s * pf
            z = gen_noise(adj, flags, True, draw)
            adj_mean = adj - rev_f
            adj = adj_mean + (torch.zeros(()) if probability_flow else G) * z
            if rec is not None:
                rec.update(p_score_adj=s)
        else:
            # solver.py:45-61 + sde.py:85-92 : noise draw first, then the score
            dt = -1.0 / sde_x.N
            z = gen_noise(x1, flags, False, draw)
            g = sde_x.diffusion(t)
            s = sfx(x1, adj, flags, t)
            drift = sde_x.drift_coeff(t) * x1 - g ** 2 * s * pf
            x_mean = x1 + drift * dt
            x = x_mean + (0.0 if probability_flow else g) * np.sqrt(-dt) * z
            if rec is not None:
                rec.update(p_score_x=s)
            dt = -1.0 / sde_adj.N
            z = gen_noise(adj, flags, True, draw)
            g = sde_adj.diffusion(t)
            s = sfa(x1, adj, flags, t)
            drift = sde_adj.drift_coeff(t) * adj - g ** 2 * s * pf
            adj_mean = adj + drift * dt
            adj = adj_mean + (0.0 if probability_flow else g) * np.sqrt(-dt) * z
            if rec is not None:
                rec.update(p_score_adj=s)
        if rec is not None:
            rec.update(x=x, adj=adj, x_mean=x_mean, adj_mean=adj_mean)
            trace.append(rec)
    n_evals = steps * (n_steps + 1)
    return (x_mean if denoise else x), (adj_mean if denoise else adj), n_evals


def s4_sample(sde_x: SDESpec, sde_adj: SDESpec, sd_x: StateDict, p_x: dict, sd_a: StateDict, p_a: dict,
              shape_x, shape_adj, flags: Tensor, draw: Draw, snr: float = 0.1, scale_eps: float = 1.0,
              denoise: bool = True, eps: float = 1e-3, trace: Optional[List[dict]] = None,
              max_iters: Optional[int] = None):
    """solver.py:200-280."""
    sfx = make_score_fn(sde_x, sd_x, p_x)
    sfa = make_score_fn(sde_adj, sd_a, p_a)
    x, adj = prior(shape_x, shape_adj, flags, draw)
    steps = sde_adj.N
    ts = torch.linspace(sde_adj.T, eps, steps)
    dt = -1.0 / steps
    x_mean, adj_mean = x, adj
    for i in range(steps if max_iters is None else min(steps, max_iters)):
        t = ts[i]
        hdt = torch.ones(()) * (dt / 2)
        sx = sfx(x, adj, flags, t)                               # solver.py:228-229
        sa = sfa(x, adj, flags, t)
        drift_x = -sde_x.diffusion(t) ** 2 * sx                 # solver.py:231-232
        drift_a = -sde_adj.diffusion(t) ** 2 * sa
        # correction (solver.py:235-258); the alpha index uses sde_x.N for both
        tidx = int((t * (sde_x.N - 1) / sde_x.T).long())
        ax = sde_x.alphas[tidx] if sde_x.kind == "VP" else torch.ones(())
        aa = sde_adj.alphas[tidx] if sde_adj.kind == "VP" else torch.ones(())
        z = gen_noise(x, flags, False, draw)
        x, _, ex = langevin_update(x, sx, z, snr, scale_eps, ax)
        z = gen_noise(adj, flags, True, draw)
        adj, _, ea = langevin_update(adj, sa, z, snr, scale_eps, aa)
        # prediction (solver.py:261-277)
        mx, sgx = sde_x.transition(t, hdt)
        ma, sga = sde_adj.transition(t, hdt)
        x = mx * x + sgx * gen_noise(x, flags, False, draw)
        adj = ma * adj + sga * gen_noise(adj, flags, True, draw)
        x = x + drift_x * dt
        adj = adj + drift_a * dt
        mx, sgx = sde_x.transition(t + hdt, hdt)
        ma, sga = sde_adj.transition(t + hdt, hdt)
        x_mean = mx * x
        adj_mean = ma * adj
        x = x_mean + sgx * gen_noise(x, flags, False, draw)
        adj = adj_mean + sga * gen_noise(adj, flags, True, draw)
        if trace is not None:
            trace.append(dict(score_x=sx, score_adj=sa, eps_x=ex, eps_adj=ea, x=x, adj=adj,
                              x_mean=x_mean, adj_mean=adj_mean))
    return (x_mean if denoise else x), (adj_mean if denoise else adj), 0


# --------------------------------------------------------------------------
# draw providers
# --------------------------------------------------------------------------
def global_rng_draw(shape):
    """Draw from the global torch CPU generator, exactly like the reference."""
    return torch.randn(*shape)


def device_rng_draw(device):
    """Draws on `device` (the reference draws its in-loop noise on the state's device, graph_utils.py:79).
    With tensors on a CUDA device every function of this module runs as plain PyTorch eager CUDA ops --
    the same ops the unmodified reference issues -- which bench.py times as the 'torch eager on the same
    GPU' baseline.  Still test/bench infrastructure only."""
    def draw(shape):
        return torch.randn(*shape, device=device)
    return draw


class RecordedDraws:
    """Replays a fixed list of raw draws (and records what a provider produced)."""

    def __init__(self, source: Optional[Draw] = None, tape: Optional[List[Tensor]] = None):
        self.source, self.tape, self.pos = source, ([] if tape is None else tape), 0

    def __call__(self, shape):
        if self.source is not None:
            t = self.source(shape)
            self.tape.append(t)
            return t
        t = self.tape[self.pos]
        self.pos += 1
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t


# --------------------------------------------------------------------------
# deterministic known-answer input recipe
# --------------------------------------------------------------------------
def kat_inputs(N: int, F: int, B: int = 4):
    idx = torch.arange(B * N * F, dtype=torch.float64)
    x = torch.sin(0.1 * idx).to(torch.float32).view(B, N, F)
    a = torch.cos(0.07 * torch.arange(B * N * N, dtype=torch.float64)).to(torch.float32).view(B, N, N)
    a = a.triu(1)
    a = a + a.transpose(-1, -2)
    counts = [N, N - 1, (2 * N) // 3, N // 2]
    counts = (counts * ((B + 3) // 4))[:B]
    flags = torch.zeros(B, N)
    for b, c in enumerate(counts):
        flags[b, :c] = 1.0
    return mask_x(x, flags), mask_adjs(a, flags), flags
