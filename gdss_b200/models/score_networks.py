"""Score networks with the reference's constructor arguments, parameter names and checkpoint
format (models/ScoreNetwork_X.py, models/ScoreNetwork_A.py, models/attention.py, models/layers.py)
whose forward pass is ONE call into the CUDA library -- these modules only hold parameters.

``state_dict()`` keys are identical to the reference's (``layers.{l}.attn.{c}.gnn_q.weight`` ...),
so ``load_state_dict`` accepts the shipped ``.pth`` checkpoints unchanged.  The forward carries no
autograd graph: training goes through ``gdss_b200.trainer.DSMTrainStep`` (hand-written backward)."""
import math

import torch
from torch import nn

from ..engine import Engine


class _GCNParams(nn.Module):
    """Parameters of DenseGCNConv (layers.py:25-47): weight [in, out] (Glorot), bias [out] (zeros)."""

    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fin, fout))
        self.bias = nn.Parameter(torch.zeros(fout))
        a = math.sqrt(6.0 / (fin + fout))
        nn.init.uniform_(self.weight, -a, a)


class _MLPParams(nn.Module):
    """Parameters of MLP (layers.py:96-133, use_bn=False): ``linears`` = num_layers nn.Linear."""

    def __init__(self, num_layers, fin, hidden, fout):
        super().__init__()
        if num_layers < 2:
            raise ValueError("only MLPs with >= 2 linears appear on the sampling path")
        dims = [fin] + [hidden] * (num_layers - 1) + [fout]
        self.linears = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(num_layers))


class _AttentionParams(nn.Module):
    """attention.py:12-23 with conv='GCN'."""

    def __init__(self, fin, attn_dim, fout):
        super().__init__()
        self.gnn_q = _GCNParams(fin, attn_dim)
        self.gnn_k = _GCNParams(fin, attn_dim)
        self.gnn_v = _GCNParams(fin, fout)


class _AttentionLayerParams(nn.Module):
    """attention.py:77-91."""

    def __init__(self, num_linears, conv_in, attn_dim, conv_out, c_in, c_out):
        super().__init__()
        self.attn = nn.ModuleList(_AttentionParams(conv_in, attn_dim, conv_out) for _ in range(c_in))
        hidden = 2 * max(c_in, c_out)
        self.mlp = _MLPParams(num_linears, 2 * c_in, hidden, c_out)
        self.multi_channel = _MLPParams(2, c_in * conv_out, hidden, conv_out)


class _CudaScoreNet(nn.Module):
    _is_adj = False

    def _engine(self, n_nodes, device):
        # validated by a content fingerprint: in-place writes that do not bump ``_version`` (``param.data.copy_`` in
        # utils/ema.py:46, :71) must re-pack the weights too
        from ..solver import _fingerprint
        key = (int(n_nodes), str(device), _fingerprint(self))
        if getattr(self, "_eng_key", None) != key:
            self._eng = Engine(None, self, n_nodes, device) if self._is_adj else Engine(self, None, n_nodes, device)
            self._eng_key = key
        return self._eng

    @torch.no_grad()
    def forward(self, x, adj, flags):
        """x [B,N,F], adj [B,N,N], flags [B,N] -> raw network output (node scores or edge scores)."""
        if not x.is_cuda:
            raise RuntimeError("gdss_b200 score networks run on CUDA tensors only (no CPU fallback)")
        if flags is None:
            flags = torch.ones(adj.shape[:2], device=adj.device)
        eng = self._engine(adj.shape[-1], x.device)
        sx, sa = eng.score(x, adj, flags, want_x=not self._is_adj, want_adj=self._is_adj, inputs_masked=False)
        return sa if self._is_adj else sx


class ScoreNetworkX(_CudaScoreNet):
    """ScoreNetwork_X.py:9-46."""

    def __init__(self, max_feat_num, depth, nhid):
        super().__init__()
        self.nfeat, self.depth, self.nhid = max_feat_num, depth, nhid
        self.layers = nn.ModuleList(_GCNParams(max_feat_num if l == 0 else nhid, nhid) for l in range(depth))
        fdim = max_feat_num + depth * nhid
        self.final = _MLPParams(3, fdim, 2 * fdim, max_feat_num)


class ScoreNetworkX_GMH(_CudaScoreNet):
    """ScoreNetwork_X.py:49-89."""

    def __init__(self, max_feat_num, depth, nhid, num_linears, c_init, c_hid, c_final, adim, num_heads=4, conv='GCN'):
        super().__init__()
        if conv != 'GCN' or num_heads != 4:
            raise NotImplementedError("only conv='GCN', num_heads=4 (every shipped checkpoint) is built")
        self.depth, self.c_init = depth, c_init
        layers = []
        for l in range(depth):
            if l == 0:
                layers.append(_AttentionLayerParams(num_linears, max_feat_num, nhid, nhid, c_init, c_hid))
            elif l == depth - 1:
                layers.append(_AttentionLayerParams(num_linears, nhid, adim, nhid, c_hid, c_final))
            else:
                layers.append(_AttentionLayerParams(num_linears, nhid, adim, nhid, c_hid, c_hid))
        self.layers = nn.ModuleList(layers)
        fdim = max_feat_num + depth * nhid
        self.final = _MLPParams(3, fdim, 2 * fdim, max_feat_num)


class ScoreNetworkA(_CudaScoreNet):
    """ScoreNetwork_A.py:101-150."""
    _is_adj = True

    def __init__(self, max_feat_num, max_node_num, nhid, num_layers, num_linears, c_init, c_hid, c_final, adim,
                 num_heads=4, conv='GCN'):
        super().__init__()
        if conv != 'GCN' or num_heads != 4:
            raise NotImplementedError("only conv='GCN', num_heads=4 (every shipped checkpoint) is built")
        self.nfeat, self.max_node_num, self.nhid, self.num_layers = max_feat_num, max_node_num, nhid, num_layers
        self.num_linears, self.c_init, self.c_hid, self.c_final, self.adim = num_linears, c_init, c_hid, c_final, adim
        layers = []
        for l in range(num_layers):
            if l == 0:
                layers.append(_AttentionLayerParams(num_linears, max_feat_num, nhid, nhid, c_init, c_hid))
            elif l == num_layers - 1:
                layers.append(_AttentionLayerParams(num_linears, nhid, adim, nhid, c_hid, c_final))
            else:
                layers.append(_AttentionLayerParams(num_linears, nhid, adim, nhid, c_hid, c_hid))
        self.layers = nn.ModuleList(layers)
        self.fdim = c_hid * (num_layers - 1) + c_final + c_init
        self.final = _MLPParams(3, self.fdim, 2 * self.fdim, 1)


def load_model(params):
    """utils/loader.py:37-48."""
    params_ = dict(params)
    model_type = params_.pop('model_type', None)
    if model_type == 'ScoreNetworkX':
        return ScoreNetworkX(**params_)
    if model_type == 'ScoreNetworkX_GMH':
        return ScoreNetworkX_GMH(**params_)
    if model_type == 'ScoreNetworkA':
        return ScoreNetworkA(**params_)
    raise ValueError(f"Model Name <{model_type}> is Unknown")


def load_model_from_ckpt(params, state_dict, device):
    """utils/loader.py:186-196 (one process per GPU: no DataParallel wrapping)."""
    model = load_model(params)
    if 'module.' in list(state_dict.keys())[0]:
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    model.load_state_dict(state_dict)
    dev = f"cuda:{device[0]}" if isinstance(device, list) else device
    return model.to(dev)
