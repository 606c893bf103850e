"""Drop-in for the forward pass of the reference's losses.py: ``get_sde_loss_fn`` (losses.py:37-81) with the same
signature and the same returned closure

    loss_fn(model_x, model_adj, x, adj) -> (loss_x, loss_adj)        # 0-dim CUDA tensors

This is what ``Trainer.train`` evaluates on the test loader every epoch under ``torch.no_grad()`` (trainer.py:84-99)
and the forward half of a training step.  The time draw and the two Gaussian draws are made with the same torch calls,
in the same order, as the reference (``torch.rand`` for t, ``torch.randn_like`` for z_x then z_adj), so under the same
seed on the same device the closure sees the reference's own random numbers; everything after that (node flags, noise
masking / symmetrisation, perturbation, both score networks, the per-graph reduction and the batch mean) runs in the
CUDA library (``gdss_dsm_loss``).  The returned tensors carry no autograd graph (this is the loss the trainer evaluates under
``no_grad``, trainer.py:84-99); the loop body trainer.py:57-79 -- both losses WITH their parameter gradients, clip, Adam, EMA -- is
``gdss_b200.trainer.DSMTrainStep``.
"""
import torch

from ._lib import check, lib, require_cuda
from .sde import sde_kind
from .solver import _engine_for


def _marginal(sde, t):
    """(mean coefficient, std) of sde.marginal_prob(., t) for a [B] time tensor, with the reference's expressions
    (sde.py:134-138 VP, :196-199 VE, :259-263 subVP)."""
    kind = sde_kind(sde)
    if kind == "VE":
        return torch.ones_like(t), sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** t
    lmc = -0.25 * t ** 2 * (sde.beta_1 - sde.beta_0) - 0.5 * t * sde.beta_0
    if kind == "VP":
        return torch.exp(lmc), torch.sqrt(1. - torch.exp(2. * lmc))
    return torch.exp(lmc), 1 - torch.exp(2. * lmc)


def get_sde_loss_fn(sde_x, sde_adj, train=True, reduce_mean=False, continuous=True, likelihood_weighting=False, eps=1e-5):
    """losses.py:37-81.  ``train`` only selects ``model.eval()`` in the reference (losses.py:8-9); the CUDA path has no
    train-mode state (the score networks have no dropout / batch norm)."""
    if not continuous:
        raise NotImplementedError("Discrete not supported")                      # losses.py:18, :28
    if likelihood_weighting:
        raise NotImplementedError("likelihood_weighting=True is not built (no shipped config uses it)")
    kx, ka = sde_kind(sde_x), sde_kind(sde_adj)

    def loss_fn(model_x, model_adj, x, adj, t=None, raw_x=None, raw_adj=None):
        """``t`` / ``raw_x`` / ``raw_adj`` inject the draws (parity tests); default = the reference's torch calls."""
        require_cuda()
        if not adj.is_cuda:
            raise RuntimeError("gdss_b200.losses works on CUDA tensors only (no CPU fallback)")
        dev = adj.device
        x = x.to(dev, torch.float32).contiguous()
        adj = adj.to(torch.float32).contiguous()
        B, N, F = x.shape
        with torch.no_grad():
            if t is None:
                t = torch.rand(adj.shape[0], device=dev) * (sde_adj.T - eps) + eps     # losses.py:47
            if raw_x is None:
                raw_x = torch.randn_like(x)                                            # gen_noise(x, ...), losses.py:50
            if raw_adj is None:
                raw_adj = torch.randn_like(adj)                                        # gen_noise(adj, ...), losses.py:55
            t = t.to(dev, torch.float32)
            mx, sx = _marginal(sde_x, t)
            ma, sa = _marginal(sde_adj, t)
            zero = torch.zeros_like(t)
            coef = torch.stack([mx, sx, zero if kx == "VE" else sx, ma, sa, zero if ka == "VE" else sa], dim=1).contiguous()
            eng = _engine_for(model_x, model_adj, N, dev)
            L = lib()
            with torch.cuda.device(dev):
                ws, need = eng.workspace(B)
                nfl = int(L.gdss_dsm_scratch_floats(eng.ctx, B))
                scratch = torch.empty(nfl, device=dev)
                out = torch.empty(2, device=dev)
                rx = raw_x.to(dev, torch.float32).contiguous()          # bound to locals: alive across the call
                ra = raw_adj.to(dev, torch.float32).contiguous()
                check(L.gdss_dsm_loss(eng.ctx, x.data_ptr(), adj.data_ptr(), coef.data_ptr(),
                                      rx.data_ptr(), ra.data_ptr(), B, int(bool(reduce_mean)),
                                      scratch.data_ptr(), nfl, out.data_ptr(), ws.data_ptr(), need,
                                      torch.cuda.current_stream(dev).cuda_stream), "gdss_dsm_loss")
        return out[0], out[1]

    return loss_fn
