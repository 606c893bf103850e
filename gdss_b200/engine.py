"""Host-side engine: packs reference state_dicts, owns the device blobs / context / workspace and
calls the C ABI (include/gdss_b200.h).  PyTorch is used only for device memory and streams."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import NetDesc, SamplerDesc, check, lib, require_cuda


def _strip(sd):
    """utils/loader.py:188-190 -- drop DataParallel's 'module.' prefix."""
    keys = list(sd.keys())
    if keys and "module." in keys[0]:
        return {k[7:]: v for k, v in sd.items()}
    return dict(sd)


def model_state(model):
    """state_dict of one of our modules, a reference nn.Module (possibly DataParallel-wrapped) or a
    plain dict of tensors."""
    if isinstance(model, dict):
        return _strip(model)
    if hasattr(model, "module") and hasattr(model.module, "state_dict"):
        model = model.module
    return _strip(model.state_dict())


def resolve_device(device):
    """'cuda' / 'cuda:1' / torch.device -> a torch.device with an explicit index (the current device for a bare 'cuda')."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def describe_network(sd, n_nodes):
    """Infer the constructor arguments (utils/loader.py:150-165 params dict) from the tensor shapes
    of a state_dict (key grammar: SURVEY.md 8b).  -> NetDesc"""
    d = NetDesc()
    keys = list(sd.keys())
    if "layers.0.weight" in sd:                                           # ScoreNetworkX
        depth = len([k for k in keys if k.startswith("layers.") and k.endswith(".weight")])
        d.model_type = _lib.NET_X
        d.max_feat_num, d.nhid = sd["layers.0.weight"].shape
        d.depth = depth
        d.num_linears, d.c_init, d.c_hid, d.c_final, d.adim, d.num_heads = 3, 1, 1, 1, d.nhid, 4
    elif "layers.0.attn.0.gnn_q.weight" in sd:
        depth = 1 + max(int(k.split(".")[1]) for k in keys if k.startswith("layers."))
        nlin = 1 + max(int(k.split(".")[4]) for k in keys if k.startswith("layers.0.mlp.linears."))
        fout = sd["final.linears.2.weight"].shape[0]
        d.max_feat_num, d.nhid = sd["layers.0.attn.0.gnn_q.weight"].shape
        is_a = fout == 1 and sd["final.linears.0.weight"].shape[1] != d.max_feat_num + depth * d.nhid
        d.model_type = _lib.NET_A if is_a else _lib.NET_X_GMH
        d.depth, d.num_linears = depth, nlin
        d.c_init = 1 + max(int(k.split(".")[3]) for k in keys if k.startswith("layers.0.attn."))
        d.c_hid = sd[f"layers.0.mlp.linears.{nlin - 1}.weight"].shape[0]
        d.c_final = sd[f"layers.{depth - 1}.mlp.linears.{nlin - 1}.weight"].shape[0]
        d.adim = sd["layers.1.attn.0.gnn_q.weight"].shape[1] if depth > 1 else d.nhid
        d.num_heads = 4                                                  # every constructor default / config
    else:
        raise ValueError("Model Name <unknown state_dict> is Unknown")   # utils/loader.py:46
    d.max_node_num = n_nodes
    return d


def state_keys(desc):
    """The state_dict keys of a network in the order gdss_pack_weights consumes them (key grammar: SURVEY.md 8b,
    models/ScoreNetwork_X.py:17-30, models/attention.py:77-91, models/layers.py:96-133)."""
    keys = []
    if desc.model_type == _lib.NET_X:
        for l in range(desc.depth):
            keys += [f"layers.{l}.weight", f"layers.{l}.bias"]
    else:
        for l in range(desc.depth):
            c_in = desc.c_init if l == 0 else desc.c_hid
            for c in range(c_in):
                for m in "qkv":
                    keys += [f"layers.{l}.attn.{c}.gnn_{m}.weight", f"layers.{l}.attn.{c}.gnn_{m}.bias"]
            for i in range(desc.num_linears):
                keys += [f"layers.{l}.mlp.linears.{i}.weight", f"layers.{l}.mlp.linears.{i}.bias"]
            for i in range(2):
                keys += [f"layers.{l}.multi_channel.linears.{i}.weight", f"layers.{l}.multi_channel.linears.{i}.bias"]
    for i in range(3):
        keys += [f"final.linears.{i}.weight", f"final.linears.{i}.bias"]
    return keys


def pack_network(sd, n_nodes, desc=None, model=None):
    """-> (NetDesc, host blob float32 tensor) via gdss_pack_weights.  Tensors are looked up BY KEY (a sorted, filtered
    or rebuilt state_dict packs correctly or fails loudly); ``num_heads`` is read from the module when it has one and
    anything but 4 is rejected (the kernels are built for the 4 heads of every shipped config)."""
    L = lib()
    sd = _strip(sd)
    desc = desc or describe_network(sd, n_nodes)
    heads = None
    if model is not None and not isinstance(model, dict):
        m = model.module if hasattr(model, "module") and hasattr(model.module, "state_dict") else model
        heads = getattr(m, "num_heads", None)
        if heads is None:
            try:
                heads = m.layers[0].attn[0].num_heads
            except Exception:
                heads = None
    if heads is not None and int(heads) != 4:
        raise _lib.GdssError(f"num_heads={heads} is not supported (the kernels are built for num_heads=4)")
    keys = state_keys(desc)
    missing = [k for k in keys if k not in sd]
    known = set(keys)
    extra = [k for k in sd if k not in known]
    if missing or extra:
        raise _lib.GdssError(f"state_dict does not match the network description: missing {missing[:3]}, unexpected {extra[:3]}")
    tensors = [sd[k].detach().to("cpu", torch.float32).contiguous() for k in keys]
    n = len(tensors)
    total = L.gdss_packed_floats(C.byref(desc), n_nodes)
    if total < 0:
        check(int(total), "gdss_packed_floats")
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    numel = (C.c_int64 * n)(*[t.numel() for t in tensors])
    blob = torch.zeros(int(total), dtype=torch.float32)
    check(L.gdss_pack_weights(C.byref(desc), n_nodes, ptrs, numel, n, blob.data_ptr(), total), "gdss_pack_weights")
    return desc, blob


class Engine:
    """Two packed score networks bound to (N, F) on one CUDA device."""

    def __init__(self, model_x, model_adj, n_nodes, device="cuda"):
        require_cuda()
        self.L = lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GdssError("the GDSS kernels run on a CUDA device only (no CPU fallback)")
        self.N = int(n_nodes)
        self.desc_x = self.desc_a = None
        self.blob_x = self.blob_a = None
        with torch.cuda.device(self.device):
            if model_x is not None:
                self.desc_x, hb = pack_network(model_state(model_x), self.N, model=model_x)
                self.blob_x = hb.to(self.device)
            if model_adj is not None:
                self.desc_a, hb = pack_network(model_state(model_adj), self.N, model=model_adj)
                self.blob_a = hb.to(self.device)
                if self.desc_a.model_type != _lib.NET_A:
                    raise ValueError("model_adj must be a ScoreNetworkA")
            ref = self.desc_x or self.desc_a
            self.F = int(ref.max_feat_num)
            self.ctx = self.L.gdss_ctx_create(self.N, self.F,
                                              C.byref(self.desc_x) if self.desc_x else None,
                                              self.blob_x.data_ptr() if self.blob_x is not None else None,
                                              C.byref(self.desc_a) if self.desc_a else None,
                                              self.blob_a.data_ptr() if self.blob_a is not None else None)
        if not self.ctx:
            raise _lib.GdssError("gdss_ctx_create failed (unsupported network configuration)")
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.L.gdss_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def workspace(self, B):
        need = int(self.L.gdss_workspace_bytes(self.ctx, B))
        if need < 0:
            check(need, "gdss_workspace_bytes")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, need

    def region(self, B, name, shape, dtype=torch.float32):
        """View of a named workspace region (parity tests look at intermediates)."""
        ws, _ = self.workspace(B)
        off = int(self.L.gdss_workspace_offset(self.ctx, B, name.encode()))
        if off < 0:
            raise KeyError(name)
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return ws[off:off + n].view(dtype).view(*shape)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def score(self, x, adj, flags, want_x=True, want_adj=True, inputs_masked=False):
        """Raw network outputs (ScoreNetworkX*.forward / ScoreNetworkA.forward)."""
        B = x.shape[0]
        x, adj, flags = [t.to(self.device, torch.float32).contiguous() for t in (x, adj, flags)]
        with torch.cuda.device(self.device):
            ws, need = self.workspace(B)
            sx = torch.empty_like(x) if (want_x and self.desc_x) else None
            sa = torch.empty_like(adj) if (want_adj and self.desc_a) else None
            check(self.L.gdss_score_forward(self.ctx, x.data_ptr(), adj.data_ptr(), flags.data_ptr(),
                                            sx.data_ptr() if sx is not None else None,
                                            sa.data_ptr() if sa is not None else None,
                                            B, int(bool(inputs_masked)), ws.data_ptr(), need, self._stream()),
                  "gdss_score_forward")
        return sx, sa

    def sample(self, sdesc, table, flags, x=None, adj=None, first_step=0, n_iters=None, seed=0, graph_offset=0,
               global_B=None, inj_x=None, inj_adj=None, trace=False, comm=None, use_graph=True):
        """Run solver steps [first_step, first_step + n_iters) -> (x, adj, x_mean, adj_mean[, trace])."""
        flags = flags.to(self.device, torch.float32).contiguous()
        B, N, F = flags.shape[0], self.N, self.F
        n_iters = table.shape[0] - first_step if n_iters is None else int(n_iters)
        table = np.ascontiguousarray(table[: first_step + n_iters], dtype=np.float32)
        init_prior = x is None
        with torch.cuda.device(self.device):
            if init_prior:
                x = torch.empty(B, N, F, device=self.device)
                adj = torch.empty(B, N, N, device=self.device)
            else:
                x = x.to(self.device, torch.float32).contiguous().clone()
                adj = adj.to(self.device, torch.float32).contiguous().clone()
            xm, am = torch.empty_like(x), torch.empty_like(adj)
            tr = torch.zeros(max(n_iters, 1), 8, device=self.device) if trace else None
            if inj_x is not None:
                inj_x = inj_x.to(self.device, torch.float32).contiguous()
                inj_adj = inj_adj.to(self.device, torch.float32).contiguous()
            ws, need = self.workspace(B)
            rc = self.L.gdss_sample(self.ctx, C.byref(sdesc), table.ctypes.data, first_step, n_iters,
                                    x.data_ptr(), adj.data_ptr(), flags.data_ptr(), xm.data_ptr(), am.data_ptr(), B,
                                    int(init_prior), int(seed) & (2 ** 64 - 1), int(graph_offset),
                                    int(global_B if global_B is not None else B),
                                    inj_x.data_ptr() if inj_x is not None else None,
                                    inj_adj.data_ptr() if inj_adj is not None else None,
                                    tr.data_ptr() if tr is not None else None,
                                    comm, int(bool(use_graph)), ws.data_ptr(), need, self._stream())
            check(rc, "gdss_sample")
        return (x, adj, xm, am, tr) if trace else (x, adj, xm, am)


    def solver_phase(self, sdesc, table, step, phases, x, adj, flags, seed=0, graph_offset=0, global_B=None, inj_x=None,
                     inj_adj=None, comm=None):
        """ONE phase of solver step `step` (gdss_solver_phase): phases = _lib.PHASE_CORRECTOR, _lib.PHASE_PREDICTOR or both.
        x / adj are updated in place (float32, contiguous, on the engine's device); inj_x / inj_adj = this step's draws
        [draws, B, N, F|N].  -> (x_mean, adj_mean, sums4) with sums4 = the batch sums of the corrector's four norms."""
        flags = flags.to(self.device, torch.float32).contiguous()
        B = flags.shape[0]
        row = np.ascontiguousarray(table[step:step + 1], dtype=np.float32)
        if not (x.is_cuda and adj.is_cuda) or x.dtype != torch.float32 or adj.dtype != torch.float32 or not x.is_contiguous() \
                or not adj.is_contiguous():
            raise ValueError("solver_phase updates x / adj in place: pass contiguous fp32 tensors on the engine's device")
        with torch.cuda.device(self.device):
            xm, am = torch.empty_like(x), torch.empty_like(adj)
            sums = torch.zeros(4, device=self.device)
            if inj_x is not None:
                inj_x = inj_x.to(self.device, torch.float32).contiguous()
                inj_adj = inj_adj.to(self.device, torch.float32).contiguous()
            ws, need = self.workspace(B)
            rc = self.L.gdss_solver_phase(self.ctx, C.byref(sdesc), row.ctypes.data, int(step), int(phases), x.data_ptr(),
                                          adj.data_ptr(), flags.data_ptr(), xm.data_ptr(), am.data_ptr(), B,
                                          int(seed) & (2 ** 64 - 1), int(graph_offset), int(global_B if global_B is not None else B),
                                          inj_x.data_ptr() if inj_x is not None else None,
                                          inj_adj.data_ptr() if inj_adj is not None else None, sums.data_ptr(), comm,
                                          ws.data_ptr(), need, self._stream())
            check(rc, "gdss_solver_phase")
        return xm, am, sums


def make_sampler_desc(sde_x, sde_adj, predictor, corrector, snr, scale_eps, eps, n_steps, probability_flow, denoise):
    from .sde import sde_params
    d = SamplerDesc()
    for dst, sde in ((d.sde_x, sde_x), (d.sde_adj, sde_adj)):
        kind, lo, hi, N = sde_params(sde)
        dst.type = {"VP": _lib.SDE_VP, "VE": _lib.SDE_VE, "subVP": _lib.SDE_SUBVP}[kind]
        dst.num_scales, dst.lo, dst.hi = N, lo, hi
    d.predictor = _lib.PRED_S4 if predictor == "S4" else (_lib.PRED_REVERSE if predictor == "Reverse" else _lib.PRED_EULER)
    d.corrector = _lib.CORR_LANGEVIN if corrector == "Langevin" else _lib.CORR_NONE
    d.snr, d.scale_eps, d.eps = float(snr), float(scale_eps), float(eps)
    d.n_steps, d.probability_flow, d.denoise = int(n_steps), int(bool(probability_flow)), int(bool(denoise))
    return d
