"""CUDA versions of the tensor helpers the sampling path uses (utils/graph_utils.py): same names,
same argument meaning, CUDA tensors only."""
import numpy as np
import torch

from ._lib import check, lib, require_cuda


def _prep(t):
    if not t.is_cuda:
        raise RuntimeError("gdss_b200.graph_utils works on CUDA tensors only (no CPU fallback)")
    return t.to(torch.float32).contiguous()


def _fresh_seed():
    """New 64-bit Philox seed per call from torch's default generator (see gdss_b200.solver.fresh_seed)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def mask_x(x, flags):
    """utils/graph_utils.py:8-12."""
    if flags is None:
        return x
    require_cuda()
    out, flags = _prep(x).clone(), _prep(flags)
    B, N, F = out.shape
    with torch.cuda.device(out.device):
        check(lib().gdss_mask_x(out.data_ptr(), flags.data_ptr(), B, N, F, _stream(out)), "gdss_mask_x")
    return out


def mask_adjs(adjs, flags):
    """utils/graph_utils.py:16-29 ([B,N,N] or [B,C,N,N])."""
    if flags is None:
        return adjs
    require_cuda()
    out, flags = _prep(adjs).clone(), _prep(flags)
    B, N = out.shape[0], out.shape[-1]
    Cn = out.shape[1] if out.dim() == 4 else 1
    with torch.cuda.device(out.device):
        check(lib().gdss_mask_adjs(out.data_ptr(), flags.data_ptr(), B, Cn, N, _stream(out)), "gdss_mask_adjs")
    return out


def gen_noise(x, flags, sym=True, raw=None, seed=None, draw_id=0, graph_offset=0):
    """utils/graph_utils.py:78-86.  ``raw`` (same shape as x) injects the N(0,1) draw; otherwise
    the library's Philox stream keyed by (seed, graph id, draw_id) is used."""
    require_cuda()
    x, flags = _prep(x), _prep(flags)
    out = torch.empty_like(x)
    B, N = x.shape[0], x.shape[1]
    F = 0 if sym else x.shape[2]
    seed = _fresh_seed() if seed is None else seed          # a new stream per call, like torch.randn_like in the reference
    raw_t = _prep(raw) if raw is not None else None         # kept alive across the call
    rawp = raw_t.data_ptr() if raw_t is not None else None
    with torch.cuda.device(x.device):
        check(lib().gdss_gen_noise(out.data_ptr(), rawp, flags.data_ptr(), B, N, F, int(bool(sym)),
                                   int(seed) & (2 ** 64 - 1), int(draw_id), int(graph_offset), _stream(x)), "gdss_gen_noise")
    return out


def pow_tensor(x, cnum):
    """utils/graph_utils.py:133-142 -> [B, cnum, N, N]."""
    require_cuda()
    x = _prep(x)
    B, N = x.shape[0], x.shape[-1]
    out = torch.empty(B, cnum, N, N, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().gdss_pow_tensor(x.data_ptr(), out.data_ptr(), B, N, cnum, _stream(x)), "gdss_pow_tensor")
    return out


def quantize(adjs, thr=0.5):
    """utils/graph_utils.py:90-92."""
    require_cuda()
    a = _prep(adjs)
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        check(lib().gdss_quantize(a.data_ptr(), out.data_ptr(), float(thr), a.numel(), _stream(a)), "gdss_quantize")
    return out


def quantize_mol(adjs):
    """utils/graph_utils.py:97-106 -> int64 numpy array with bond orders {0,1,2,3}."""
    require_cuda()
    a = _prep(adjs if torch.is_tensor(adjs) else torch.as_tensor(np.asarray(adjs)).cuda())
    out = torch.empty(a.shape, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        check(lib().gdss_quantize_mol(a.data_ptr(), out.data_ptr(), a.numel(), _stream(a)), "gdss_quantize_mol")
    return out.cpu().numpy()


def decode_mol(x, adj, one_hot=True):
    """sampler.py:131-138 on the device: quantize_mol -> bond-class remap (0,1,2,3 -> 3,0,1,2) and x > 0.5 with the
    virtual-atom channel appended.  Returns (x [B,N,F+1], adj) where adj is the reference's one-hot [B,4,N,N] int64
    (one_hot=True) or the compact int8 class index [B,N,N]; the fp32 state never leaves the device."""
    require_cuda()
    x, adj = _prep(x), _prep(adj)
    B, N, F = x.shape
    atom = torch.empty(B, N, F + 1, dtype=torch.int8, device=x.device)
    bond = torch.empty(B, N, N, dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().gdss_decode_mol(x.data_ptr(), adj.data_ptr(), atom.data_ptr(), bond.data_ptr(), B, N, F, _stream(x)),
              "gdss_decode_mol")
    if not one_hot:
        return atom, bond
    return atom.to(torch.int64), torch.nn.functional.one_hot(bond.to(torch.int64), num_classes=4).permute(0, 3, 1, 2)


def init_flags_device(node_counts, max_node_num, batch_size, device="cuda", seed=None, graph_offset=0):
    """utils/graph_utils.py:66-74 without the host round trip: ``node_counts`` = the node numbers of the training
    graphs (any iterable of ints); flags [batch_size, max_node_num] are drawn on the device from their histogram."""
    require_cuda()
    N = int(max_node_num)
    hist = np.bincount(np.clip(np.asarray(list(node_counts), dtype=np.int64), 0, N), minlength=N + 1).astype(np.float64)
    cdf = torch.from_numpy(np.cumsum(hist).astype(np.float32)).to(device)
    flags = torch.empty(int(batch_size), N, device=device)
    seed = _fresh_seed() if seed is None else seed
    with torch.cuda.device(flags.device):
        check(lib().gdss_init_flags(cdf.data_ptr(), flags.data_ptr(), int(batch_size), N, int(seed) & (2 ** 64 - 1),
                                    int(graph_offset), _stream(flags)), "gdss_init_flags")
    return flags


def node_flags(adj, eps=1e-5):
    """utils/graph_utils.py:33-39 (host-side glue, torch)."""
    flags = torch.abs(adj).sum(-1).gt(eps).to(dtype=torch.float32)
    if len(flags.shape) == 3:
        flags = flags[:, 0, :]
    return flags
