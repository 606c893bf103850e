"""Manual reverse sweep of the score networks -- TEST INFRASTRUCTURE ONLY (the oracle of the backward kernels that the next
round builds; nothing in the product path imports it).

Every function is the hand-derived gradient of the forward restatement of the same name in oracle/gdss_oracle.py, written
with explicit tensor algebra and NO torch.autograd, in the form a per-graph CUDA kernel implements it.  They are checked
against torch.autograd over the forward oracle (tests/test_oracle_golden.py), which is itself pinned to the reference's
autograd (tests/golden/dsm_grad_*.npz).  Citations are the reference lines of the forward being differentiated.

Conventions: d<name> is dLoss/d<name>; grads is a dict  state_dict key -> gradient tensor (accumulated with +=).
"""
import math
from typing import Dict, List, Tuple

import torch

from . import gdss_oracle as O

Tensor = torch.Tensor


def _acc(grads: Dict[str, Tensor], key: str, g: Tensor):
    grads[key] = grads[key] + g if key in grads else g


# ---------------------------------------------------------------------------------------------------------------------------
# DenseGCNConv (models/layers.py:49-88):  A~ = adj with unit diagonal, deg = clamp(rowsum A~, 1), d = deg^-1/2,
#   A^ = d_i A~_ij d_j,  Y = A^ (x W) + b
# ---------------------------------------------------------------------------------------------------------------------------
def dense_gcn_forward(x, adj, W, b):
    n = adj.shape[-1]
    eye = torch.eye(n, dtype=adj.dtype)
    at = adj * (1 - eye) + eye
    deg_raw = at.sum(-1)
    d = deg_raw.clamp(min=1).pow(-0.5)
    ah = d.unsqueeze(-1) * at * d.unsqueeze(-2)
    xw = x @ W
    return ah @ xw + b, (x, at, deg_raw, d, ah, xw, W)


def dense_gcn_backward(saved, dY) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (dx [B,N,Fin], dadj [B,N,N] (zero diagonal), dW [Fin,Fout], db [Fout])."""
    x, at, deg_raw, d, ah, xw, W = saved
    n = at.shape[-1]
    db = dY.sum(dim=(0, 1))
    dxw = ah.transpose(-1, -2) @ dY                       # [B,N,Fout]
    dW = (x.transpose(-1, -2) @ dxw).sum(0)
    dx = dxw @ W.t()
    G = dY @ xw.transpose(-1, -2)                         # dA^_ij = sum_o dY_io (xW)_jo
    dat = d.unsqueeze(-1) * G * d.unsqueeze(-2)           # through A^_ij = d_i A~_ij d_j  (A~ held fixed)
    dd = (G * at * d.unsqueeze(-2)).sum(-1) + (G * at * d.unsqueeze(-1)).sum(-2)   # d as the row factor + as the column factor
    ddeg = -0.5 * d.pow(3) * dd * (deg_raw >= 1).to(dd.dtype)                      # d = clamp(deg, 1)^-1/2
    dat = dat + ddeg.unsqueeze(-1)                        # deg_i = sum_j A~_ij
    eye = torch.eye(n, dtype=at.dtype)
    return dx, dat * (1 - eye), dW, db                    # the diagonal of A~ is the constant 1


# ---------------------------------------------------------------------------------------------------------------------------
# MLP (models/layers.py:135-153): Linear, ELU, ..., Linear  (nn.Linear weights are [out, in])
# ---------------------------------------------------------------------------------------------------------------------------
def mlp_forward(sd, prefix, h, num_layers):
    saved = []
    for i in range(num_layers):
        Wt, b = sd[f"{prefix}.linears.{i}.weight"], sd[f"{prefix}.linears.{i}.bias"]
        pre = h @ Wt.t() + b
        saved.append((h, pre))
        h = torch.where(pre > 0, pre, torch.expm1(pre)) if i < num_layers - 1 else pre
    return h, saved


def mlp_backward(sd, prefix, saved, dout, grads) -> Tensor:
    num_layers = len(saved)
    g = dout
    for i in reversed(range(num_layers)):
        h, pre = saved[i]
        if i < num_layers - 1:
            g = g * torch.where(pre > 0, torch.ones_like(pre), torch.exp(pre))      # ELU'
        Wt = sd[f"{prefix}.linears.{i}.weight"]
        g2, h2 = g.reshape(-1, g.shape[-1]), h.reshape(-1, h.shape[-1])
        _acc(grads, f"{prefix}.linears.{i}.weight", g2.t() @ h2)
        _acc(grads, f"{prefix}.linears.{i}.bias", g2.sum(0))
        g = g @ Wt
    return g


# ---------------------------------------------------------------------------------------------------------------------------
# Attention (models/attention.py:25-51): Q, K, V = GCNs; S_h = Q_h K_h^T / sqrt(out_dim); T = mean_h tanh S_h; M = (T + T^T) / 2
# ---------------------------------------------------------------------------------------------------------------------------
def attention_forward(sd, prefix, x, adj_c, heads):
    q, sq = dense_gcn_forward(x, adj_c, sd[f"{prefix}.gnn_q.weight"], sd[f"{prefix}.gnn_q.bias"])
    k, sk = dense_gcn_forward(x, adj_c, sd[f"{prefix}.gnn_k.weight"], sd[f"{prefix}.gnn_k.bias"])
    v, sv = dense_gcn_forward(x, adj_c, sd[f"{prefix}.gnn_v.weight"], sd[f"{prefix}.gnn_v.bias"])
    B, n, ad = q.shape
    hd, scale = ad // heads, 1.0 / math.sqrt(v.shape[-1])
    qh = q.view(B, n, heads, hd).permute(2, 0, 1, 3)
    kh = k.view(B, n, heads, hd).permute(2, 0, 1, 3)
    th = torch.tanh(qh @ kh.transpose(-1, -2) * scale)    # [H,B,N,N]
    t = th.mean(0)
    return v, (t + t.transpose(-1, -2)) / 2, (sq, sk, sv, qh, kh, th, scale, heads)


def attention_backward(sd, prefix, saved, dV, dM, grads) -> Tuple[Tensor, Tensor]:
    """-> (dx, dadj_c)."""
    sq, sk, sv, qh, kh, th, scale, heads = saved
    dT = (dM + dM.transpose(-1, -2)) / 2
    dS = (dT / heads).unsqueeze(0) * (1 - th * th) * scale            # [H,B,N,N]
    dqh = dS @ kh                                                     # dQ_h = dS_h K_h
    dkh = dS.transpose(-1, -2) @ qh                                   # dK_h = dS_h^T Q_h
    B, n = dV.shape[0], dV.shape[1]
    dq = dqh.permute(1, 2, 0, 3).reshape(B, n, -1)
    dk = dkh.permute(1, 2, 0, 3).reshape(B, n, -1)
    dx, dadj = 0, 0
    for name, s_, g in (("q", sq, dq), ("k", sk, dk), ("v", sv, dV)):
        gx, ga, gW, gb = dense_gcn_backward(s_, g)
        dx, dadj = dx + gx, dadj + ga
        _acc(grads, f"{prefix}.gnn_{name}.weight", gW)
        _acc(grads, f"{prefix}.gnn_{name}.bias", gb)
    return dx, dadj


# ---------------------------------------------------------------------------------------------------------------------------
# AttentionLayer (models/attention.py:93-116)
# ---------------------------------------------------------------------------------------------------------------------------
def attention_layer_forward(sd, prefix, x, adjc, flags, num_linears, heads):
    C = adjc.shape[1]
    vs, ms, sat = [], [], []
    for c in range(C):
        v, m, s_ = attention_forward(sd, f"{prefix}.attn.{c}", x, adjc[:, c], heads)
        vs.append(v); ms.append(m); sat.append(s_)
    hn, s_node = mlp_forward(sd, f"{prefix}.multi_channel", torch.cat(vs, -1), 2)
    x_out = torch.tanh(O.mask_x(hn, flags))
    edge_in = torch.cat([torch.stack(ms, -1), adjc.permute(0, 2, 3, 1)], -1)         # [B,N,N,2C]
    s, s_edge = mlp_forward(sd, f"{prefix}.mlp", edge_in, num_linears)               # [B,N,N,Co]
    s = s.permute(0, 3, 1, 2)
    adj_out = O.mask_adjs(s + s.transpose(-1, -2), flags)
    return x_out, adj_out, (sat, s_node, s_edge, x_out, flags, C, vs[0].shape[-1])


def attention_layer_backward(sd, prefix, saved, dx_out, dadj_out, grads) -> Tuple[Tensor, Tensor]:
    """dx_out may be None (the last layer of ScoreNetworkA drops its node output), dadj_out may be None (the last layer of
    ScoreNetworkX_GMH drops its channels) -> (dx, dadjc)."""
    sat, s_node, s_edge, x_out, flags, C, h = saved
    B, n = x_out.shape[0], x_out.shape[1]
    dVs: List = [None] * C
    if dx_out is not None:
        dpre = O.mask_x(dx_out * (1 - x_out * x_out), flags)                          # tanh', mask_x
        dcat = mlp_backward(sd, f"{prefix}.multi_channel", s_node, dpre, grads)       # [B,N,C*h]
        dVs = [dcat[..., c * h:(c + 1) * h] for c in range(C)]
    dM: List = [None] * C
    dadjc = torch.zeros(B, C, n, n)
    if dadj_out is not None:
        g = O.mask_adjs(dadj_out, flags)
        g = (g + g.transpose(-1, -2)).permute(0, 2, 3, 1)                             # d(s) of s + s^T, as [B,N,N,Co]
        dein = mlp_backward(sd, f"{prefix}.mlp", s_edge, g, grads)                    # [B,N,N,2C]
        dM = [dein[..., c] for c in range(C)]
        dadjc = dadjc + dein[..., C:].permute(0, 3, 1, 2)
    dx = 0
    for c in range(C):
        dv = dVs[c] if dVs[c] is not None else torch.zeros(B, n, h)
        dm = dM[c] if dM[c] is not None else torch.zeros(B, n, n)
        gx, ga = attention_backward(sd, f"{prefix}.attn.{c}", sat[c], dv, dm, grads)
        dx = dx + gx
        dadjc[:, c] = dadjc[:, c] + ga
    return dx, dadjc


# ---------------------------------------------------------------------------------------------------------------------------
# networks: forward with saved activations, backward from d(output) -> grads of every parameter the loss reaches
# ---------------------------------------------------------------------------------------------------------------------------
def network_grads(sd, p: dict, x, adj, flags, dout) -> Dict[str, Tensor]:
    """Gradients of sum(out * dout) with respect to the parameters of one score network (utils/loader.py:37-48 dispatch),
    out = run_model(sd, p, x, adj, flags).  Parameters the output does not depend on are absent from the result."""
    kind, grads, heads = p["model_type"], {}, p.get("num_heads", 4)
    n = adj.shape[-1]
    if kind == "ScoreNetworkX":                                                       # ScoreNetwork_X.py:32-46
        feats, saves, h = [x], [], x
        for l in range(p["depth"]):
            pre, s_ = dense_gcn_forward(h, adj, sd[f"layers.{l}.weight"], sd[f"layers.{l}.bias"])
            h = torch.tanh(pre)
            saves.append((s_, h)); feats.append(h)
        _, s_fin = mlp_forward(sd, "final", torch.cat(feats, -1), 3)
        dcat = mlp_backward(sd, "final", s_fin, O.mask_x(dout, flags), grads)
        widths = [f.shape[-1] for f in feats]
        offs = [sum(widths[:i]) for i in range(len(widths))]
        dh = 0
        for l in reversed(range(p["depth"])):
            s_, hl = saves[l]
            g = dcat[..., offs[l + 1]:offs[l + 1] + widths[l + 1]] + dh
            gx, _, gW, gb = dense_gcn_backward(s_, g * (1 - hl * hl))
            _acc(grads, f"layers.{l}.weight", gW); _acc(grads, f"layers.{l}.bias", gb)
            dh = gx
        return grads
    adjc = O.pow_stack(adj, p["c_init"])
    if kind == "ScoreNetworkX_GMH":                                                   # ScoreNetwork_X.py:75-89
        feats, saves, h = [x], [], x
        for l in range(p["depth"]):
            h1, adjc, s_ = attention_layer_forward(sd, f"layers.{l}", h, adjc, flags, p["num_linears"], heads)
            h = torch.tanh(h1)
            saves.append((s_, h)); feats.append(h)
        _, s_fin = mlp_forward(sd, "final", torch.cat(feats, -1), 3)
        dcat = mlp_backward(sd, "final", s_fin, O.mask_x(dout, flags), grads)
        widths = [f.shape[-1] for f in feats]
        offs = [sum(widths[:i]) for i in range(len(widths))]
        dh, dadjc = 0, None
        for l in reversed(range(p["depth"])):
            s_, hl = saves[l]
            g = (dcat[..., offs[l + 1]:offs[l + 1] + widths[l + 1]] + dh) * (1 - hl * hl)     # the extra tanh (:81)
            dh, dadjc = attention_layer_backward(sd, f"layers.{l}", s_, g, dadjc, grads)
        return grads
    if kind == "ScoreNetworkA":                                                       # ScoreNetwork_A.py:131-150
        stacks, saves, h = [adjc], [], x
        for l in range(p["num_layers"]):
            h, adjc, s_ = attention_layer_forward(sd, f"layers.{l}", h, adjc, flags, p["num_linears"], heads)
            saves.append(s_); stacks.append(adjc)
        cat = torch.cat(stacks, 1).permute(0, 2, 3, 1)
        _, s_fin = mlp_forward(sd, "final", cat, 3)
        g = O.mask_adjs(dout, flags) * (1.0 - torch.eye(n)).unsqueeze(0)                # zero diagonal, mask_adjs
        dcat = mlp_backward(sd, "final", s_fin, g.unsqueeze(-1), grads).permute(0, 3, 1, 2)   # [B,fdim,N,N]
        widths = [s.shape[1] for s in stacks]
        offs = [sum(widths[:i]) for i in range(len(widths))]
        dh, dnext = None, 0
        for l in reversed(range(p["num_layers"])):
            dadj_out = dcat[:, offs[l + 1]:offs[l + 1] + widths[l + 1]] + dnext
            dh, dnext = attention_layer_backward(sd, f"layers.{l}", saves[l], dh, dadj_out, grads)
        return grads
    raise ValueError(f"Model Name <{kind}> is Unknown")


# ---------------------------------------------------------------------------------------------------------------------------
# DSM losses (losses.py:37-81, reduce_mean = False, likelihood_weighting = False): loss = mean_b 0.5 sum (score std + z)^2,
# score = -net / std (VP, subVP: losses.py:19) or net (VE: losses.py:26)
# ---------------------------------------------------------------------------------------------------------------------------
def dsm_param_grads(sde_x, sde_adj, sd_x, p_x, sd_a, p_a, x, adj, t, raw_x, raw_adj):
    """-> (loss_x, loss_adj, grads_x, grads_adj): both losses and the gradients of each with respect to its own network's
    parameters (loss_x does not depend on the adjacency network and vice versa: trainer.py:62-65 backpropagates both)."""
    def marginal(sde, y):                                    # sde.py:134-138 / :196-199 / :259-263
        if sde.kind == "VE":
            return y, sde.lo * (sde.hi / sde.lo) ** t
        lmc = -0.25 * t ** 2 * (sde.hi - sde.lo) - 0.5 * t * sde.lo
        mean = torch.exp(lmc).view((-1,) + (1,) * (y.dim() - 1)) * y
        return mean, (torch.sqrt(1. - torch.exp(2. * lmc)) if sde.kind == "VP" else 1 - torch.exp(2. * lmc))

    B = x.shape[0]
    flags = O.node_flags(adj)
    z_x = O.mask_x(raw_x, flags)
    mean_x, std_x = marginal(sde_x, x)
    pert_x = O.mask_x(mean_x + std_x[:, None, None] * z_x, flags)
    z_a = O.sym_noise_from_raw(raw_adj, flags)
    mean_a, std_a = marginal(sde_adj, adj)
    pert_a = O.mask_adjs(mean_a + std_a[:, None, None] * z_a, flags)
    out = []
    for sde, sd, p, z, std in ((sde_x, sd_x, p_x, z_x, std_x), (sde_adj, sd_a, p_a, z_a, std_a)):
        net = O.run_model(sd, p, pert_x, pert_a, flags)
        s3 = std[:, None, None]
        score = net if sde.kind == "VE" else -net / s3
        resid = score * s3 + z
        loss = (0.5 * resid.reshape(B, -1).pow(2).sum(-1)).mean()
        dnet = resid * (s3 if sde.kind == "VE" else -torch.ones_like(s3)) / B       # d loss / d net
        out.append((loss, network_grads(sd, p, pert_x, pert_a, flags, dnet)))
    return out[0][0], out[1][0], out[0][1], out[1][1]
