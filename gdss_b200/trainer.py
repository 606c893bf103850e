"""The reference's training step (trainer.py:57-79) on one GPU per process:

    optimizer.zero_grad(); loss_x, loss_adj = loss_fn(model_x, model_adj, x, adj)
    loss_x.backward(); loss_adj.backward()
    clip_grad_norm_(model_x.parameters(), grad_norm); clip_grad_norm_(model_adj.parameters(), grad_norm)   # per network
    optimizer_x.step(); optimizer_adj.step(); ema_x.update(...); ema_adj.update(...)

``DSMTrainStep.step(x, adj)`` is that block: both DSM losses and their parameter gradients come from ONE library call
(``gdss_dsm_loss_backward``: forward with saved activations + the hand-derived reverse sweep of both score networks), the
data-parallel gradient exchange is one NCCL all-reduce per network (``gdss_nccl_allreduce``; replaces nn.DataParallel), and
clip + Adam + EMA are the fused ``gdss_adam_ema_step``.  Parameters live as one flat fp32 vector per network in state_dict
order (what ``torch.optim.Adam(model.parameters())`` walks); the packed blobs the kernels read are refreshed from it with one
gather per step.  The random draws (t, z_x, z_adj) are made with the reference's own torch calls in its order
(losses.py:47-55) unless injected.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, require_cuda
from .engine import describe_network, model_state, pack_network, resolve_device, state_keys
from .losses import _marginal
from .optim import FusedAdamEMA, allreduce_grads, flatten_state, unflatten_state
from .sde import sde_kind


def _blob_index_map(desc, sd, n_nodes):
    """blob position -> flat state_dict index (or -1 for padding): pack a state_dict whose tensors hold their own flat
    indices (+1) through gdss_pack_weights."""
    keys = state_keys(desc)
    off, probe = 0, {}
    for k in keys:
        n = sd[k].numel()
        probe[k] = (torch.arange(off, off + n, dtype=torch.float64) + 1).to(torch.float32).view(sd[k].shape)
        off += n
    if off >= 2 ** 24:
        raise _lib.GdssError("network too large for the fp32 index probe")
    _, blob = pack_network(probe, n_nodes, desc=desc)
    return blob.to(torch.int64) - 1, off


def unreached_mask(desc, sd):
    """uint8 vector over the flat parameters: 0 where the loss never reaches a parameter (grad None in torch, skipped by
    torch.optim.Adam): the last AttentionLayer's V convolutions + node MLP of ScoreNetworkA (its x output is dropped,
    ScoreNetwork_A.py:139-141) and the last layer's Q / K convolutions + per-edge MLP of ScoreNetworkX_GMH (its adjacency
    output is dropped, ScoreNetwork_X.py:79-84)."""
    parts = []
    last = f"layers.{desc.depth - 1}."
    for k in state_keys(desc):
        dead = False
        if k.startswith(last):
            if desc.model_type == _lib.NET_A:
                dead = ".gnn_v." in k or ".multi_channel." in k
            elif desc.model_type == _lib.NET_X_GMH:
                dead = ".gnn_q." in k or ".gnn_k." in k or ".mlp." in k
        parts.append(torch.full((sd[k].numel(),), 0 if dead else 1, dtype=torch.uint8))
    return torch.cat(parts)


class _Net:
    """Flat parameters + packed blob + optimizer state of one score network."""

    def __init__(self, sd, n_nodes, dev, lr, weight_decay, grad_norm, ema):
        sd = model_state(sd)
        self.desc = describe_network(sd, n_nodes)
        keys = state_keys(self.desc)
        self.sd_order = {k: sd[k] for k in keys}
        flat, self.metas = flatten_state(self.sd_order)
        self.flat = flat.to(dev).contiguous()
        perm, n = _blob_index_map(self.desc, self.sd_order, n_nodes)
        assert n == self.flat.numel()
        self.perm = perm.to(dev)
        self.valid = self.perm >= 0
        self.src = self.perm.clamp_min(0)
        self.valid_idx = torch.nonzero(self.valid).flatten()        # blob positions that hold a parameter ...
        self.dst_idx = self.perm[self.valid_idx]                    # ... and the flat index each of them maps to
        self.blob = torch.zeros(self.perm.numel(), device=dev)
        self.gblob = torch.zeros_like(self.blob)
        self.grad = torch.zeros_like(self.flat)
        self.opt = FusedAdamEMA(self.flat, lr=lr, weight_decay=weight_decay, grad_norm=grad_norm, ema_decay=ema,
                                active=unreached_mask(self.desc, self.sd_order))
        self.refresh_blob()

    def refresh_blob(self):
        self.blob.copy_(torch.where(self.valid, self.flat[self.src], torch.zeros((), device=self.flat.device)))

    def collect_grad(self):
        # (index tensors precomputed: no data-dependent shapes, so the step can be captured into a CUDA graph)
        self.grad.zero_()
        self.grad.index_copy_(0, self.dst_idx, self.gblob.index_select(0, self.valid_idx))
        return self.grad

    def state_dict(self, ema=False):
        src = self.opt.shadow if (ema and self.opt.shadow is not None) else self.flat
        return {k: v.clone() for k, v in unflatten_state(src, self.metas).items()}


class DSMTrainStep:
    """One optimisation step of trainer.py:57-79 per call.  ``model_x`` / ``model_adj``: nn.Modules (ours or the reference's)
    or state_dicts; hyper-parameters as config.train (utils/loader.py:51-64): lr, weight_decay, grad_norm, ema."""

    def __init__(self, model_x, model_adj, n_nodes, sde_x, sde_adj, lr=1e-3, weight_decay=0.0, grad_norm=1.0, ema=0.999,
                 reduce_mean=False, eps=1e-5, device="cuda", shard=None, use_graph=False):
        require_cuda()
        self.L = lib()
        self.dev = resolve_device(device)
        self.N = int(n_nodes)
        self.sde_x, self.sde_adj, self.eps, self.reduce_mean = sde_x, sde_adj, eps, bool(reduce_mean)
        self.kx, self.ka = sde_kind(sde_x), sde_kind(sde_adj)
        self.shard = shard
        with torch.cuda.device(self.dev):
            self.x = _Net(model_x, self.N, self.dev, lr, weight_decay, grad_norm, ema)
            self.a = _Net(model_adj, self.N, self.dev, lr, weight_decay, grad_norm, ema)
            if self.a.desc.model_type != _lib.NET_A:
                raise ValueError("model_adj must be a ScoreNetworkA")
            self.F = int(self.x.desc.max_feat_num)
            self.ctx = self.L.gdss_ctx_create(self.N, self.F, C.byref(self.x.desc), self.x.blob.data_ptr(),
                                              C.byref(self.a.desc), self.a.blob.data_ptr())
        if not self.ctx:
            raise _lib.GdssError("gdss_ctx_create failed (unsupported network configuration)")
        self._ws = None
        self.loss = torch.zeros(2, device=self.dev)
        # use_graph: the whole step (marginals, gdss_dsm_loss_backward, gradient gather, NCCL all-reduce, clip / Adam / EMA, blob
        # refresh) is captured once per batch shape into a CUDA graph and replayed; inputs go through static buffers and the
        # optimizer's per-step scalars through device memory (gdss_adam_ema_step_dev).  What it buys: the ~330 launches of a
        # step are launch-bound on small per-GPU batches (data-parallel training of B = 1024 on 8 GPUs)
        # Data-parallel runs launch eagerly unless GDSS_TRAIN_GRAPH_DP=1: with the NCCL all-reduce inside a global-mode capture and
        # no eager first step the first 8-rank attempt hung (NCCL sets its connections up lazily on the first collective).  The
        # capture now happens on the SECOND call of a batch shape, after one eager step, in thread-local capture mode (what the
        # sampler's own captured step does).  Measured on 2 ranks: 152 vs 138.5 steps/s, results as the eager run's -- but the
        # process then did not exit (captured graphs referencing the communicator at teardown; `close()` is the untested
        # remedy), hence opt-in for more than one rank.
        import os
        dp_ok = shard is None or shard.world == 1 or os.environ.get("GDSS_TRAIN_GRAPH_DP", "") == "1"
        self.use_graph = bool(use_graph) and dp_ok
        self._graphs = {}
        self.captured_launches = 0

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.L.gdss_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def _workspace(self, B):
        need = int(self.L.gdss_train_workspace_bytes(self.ctx, B))
        if need < 0:
            check(need, "gdss_train_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        return self._ws, need

    def loss_and_grads(self, x, adj, t=None, raw_x=None, raw_adj=None):
        """-> (loss tensor [2] on the device, flat grad_x, flat grad_adj) of THIS rank's batch (no exchange, no update)."""
        dev = self.dev
        x = x.to(dev, torch.float32).contiguous()
        adj = adj.to(dev, torch.float32).contiguous()
        B = x.shape[0]
        with torch.no_grad(), torch.cuda.device(dev):
            if t is None:
                t = torch.rand(B, device=dev) * (self.sde_adj.T - self.eps) + self.eps       # losses.py:47
            if raw_x is None:
                raw_x = torch.randn_like(x)                                                   # losses.py:50
            if raw_adj is None:
                raw_adj = torch.randn_like(adj)                                               # losses.py:55
            t = t.to(dev, torch.float32)
            rx = raw_x.to(dev, torch.float32).contiguous()
            ra = raw_adj.to(dev, torch.float32).contiguous()
            mx, sx = _marginal(self.sde_x, t)
            ma, sa = _marginal(self.sde_adj, t)
            zero = torch.zeros_like(t)
            coef = torch.stack([mx, sx, zero if self.kx == "VE" else sx, ma, sa, zero if self.ka == "VE" else sa], dim=1).contiguous()
            ws, need = self._workspace(B)
            check(self.L.gdss_dsm_loss_backward(self.ctx, x.data_ptr(), adj.data_ptr(), coef.data_ptr(), rx.data_ptr(), ra.data_ptr(),
                                                B, int(self.reduce_mean), self.x.gblob.data_ptr(), self.a.gblob.data_ptr(),
                                                self.loss.data_ptr(), ws.data_ptr(), need,
                                                torch.cuda.current_stream(dev).cuda_stream), "gdss_dsm_loss_backward")
            return self.loss, self.x.collect_grad(), self.a.collect_grad()

    def _eager_step(self, x, adj, t=None, raw_x=None, raw_adj=None):
        loss, gx, ga = self.loss_and_grads(x, adj, t, raw_x, raw_adj)
        world = 1 if self.shard is None else self.shard.world
        with torch.no_grad(), torch.cuda.device(self.dev):
            norms = []
            for net, g in ((self.x, gx), (self.a, ga)):
                allreduce_grads(self.shard, g)
                norms.append(net.opt.step(g, grad_scale=1.0 / world))
                net.refresh_blob()
        lx, la = loss[0].clone(), loss[1].clone()
        return lx, la, norms[0], norms[1]

    def _graph_step(self, x, adj, t, raw_x, raw_adj):
        dev = self.dev
        B = x.shape[0]
        key = (B, t is None, raw_x is None, raw_adj is None)
        world = 1 if self.shard is None else self.shard.world
        g = self._graphs.get(key)
        if g is None:
            # first call of this batch shape: eager (sets the kernels' attributes, warms the NCCL communicator up)
            self._graphs[key] = {"pending": True}
            return self._eager_step(x, adj, t, raw_x, raw_adj)
        if g.get("pending"):
            g = {"x": torch.empty(B, self.N, self.F, device=dev), "adj": torch.empty(B, self.N, self.N, device=dev),
                 "t": None if t is None else torch.empty(B, device=dev),
                 "rx": None if raw_x is None else torch.empty(B, self.N, self.F, device=dev),
                 "ra": None if raw_adj is None else torch.empty(B, self.N, self.N, device=dev),
                 "out": torch.zeros(4, device=dev)}
            self._workspace(B)
            for net in (self.x, self.a):
                if not hasattr(net.opt, "consts"):
                    net.opt.consts = torch.zeros(4, device=dev)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            before = self.L.gdss_launch_count()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                loss, gx, ga = self.loss_and_grads(g["x"], g["adj"], g["t"], g["rx"], g["ra"])
                norms = []
                for net, gr in ((self.x, gx), (self.a, ga)):
                    allreduce_grads(self.shard, gr)
                    norms.append(net.opt.enqueue_dev(gr, grad_scale=1.0 / world))
                    net.refresh_blob()
                g["out"][:2].copy_(loss)
                for k, nrm in enumerate(norms):
                    if nrm is not None:
                        g["out"][2 + k:3 + k].copy_(nrm)
            g["graph"] = graph
            self.captured_launches = int(self.L.gdss_launch_count() - before)
            self._graphs[key] = g
        with torch.no_grad(), torch.cuda.device(dev):
            g["x"].copy_(x, non_blocking=True)
            g["adj"].copy_(adj, non_blocking=True)
            for name, src in (("t", t), ("rx", raw_x), ("ra", raw_adj)):
                if src is not None:
                    g[name].copy_(src, non_blocking=True)
            for net in (self.x, self.a):
                net.opt.advance()
            g["graph"].replay()
            out = g["out"].clone()
        clip = self.x.opt.grad_norm > 0
        return out[0], out[1], (out[2] if clip else None), (out[3] if self.a.opt.grad_norm > 0 else None)

    def step(self, x, adj, t=None, raw_x=None, raw_adj=None):
        """trainer.py:57-79 for one batch -> (loss_x, loss_adj, grad_norm_x, grad_norm_adj) as device scalars (this rank's
        losses; the gradient norms are those of the global, batch-mean gradient)."""
        if self.use_graph:
            return self._graph_step(x, adj, t, raw_x, raw_adj)
        return self._eager_step(x, adj, t, raw_x, raw_adj)

    def close(self):
        """Drop the captured graphs (they reference the NCCL communicator of a data-parallel run: destroy them before it)."""
        self._graphs.clear()
        torch.cuda.synchronize(self.dev)

    def state_dicts(self, ema=False):
        return self.x.state_dict(ema), self.a.state_dict(ema)
