"""Drop-in for the reference's solver.py: ``get_pc_sampler`` (solver.py:153-196) and ``S4_solver``
(solver.py:200-280) with the same signatures and the same returned closure

    sampler(model_x, model_adj, init_flags) -> (x, adj, n_evals)

The closure packs the two score networks (ours or the reference's nn.Modules, DataParallel-wrapped
or not), and runs the whole reverse-diffusion loop on the GPU through gdss_sample (one CUDA graph
replayed per step).  Noise comes from the library's counter-based Philox stream keyed by
(seed, global graph id, step, draw); ``torch.initial_seed()`` at call time is the seed, so
``load_seed`` (utils/loader.py:15-26) keeps its meaning.

Multi-GPU (one process per GPU): pass ``shard=ShardInfo(...)`` (gdss_b200.dist) -- the batch is
split contiguously, the four Langevin norm sums are all-reduced every step (solver.py:130-132
takes the mean over the WHOLE batch) and the result is gathered on every rank.
"""
import torch

from .engine import Engine, make_sampler_desc
from .schedule import build_table

_ENGINES = {}


def _engine_for(model_x, model_adj, n_nodes, device):
    key = (id(model_x), id(model_adj), int(n_nodes), str(device))
    ent = _ENGINES.get(key)
    ver = tuple(int(p._version) for m in (model_x, model_adj) if hasattr(m, "parameters") for p in m.parameters())
    if ent is None or ent[1] != ver:
        if len(_ENGINES) > 8:
            _ENGINES.clear()
        ent = (Engine(model_x, model_adj, n_nodes, device), ver)
        _ENGINES[key] = ent
    return ent[0]


def _make(sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps, probability_flow,
          continuous, denoise, eps, device, s4, shard=None, use_graph=True, max_iters=None):
    if s4:
        predictor = "S4"
    elif predictor not in ("Euler", "Reverse"):
        predictor = "Euler"                      # solver.py:163: anything else silently means Euler
    if corrector != "Langevin":
        corrector = "None"                       # solver.py:164
    if not s4 and corrector == "Langevin" and int(n_steps) != 1:
        raise NotImplementedError("only n_steps=1 (every shipped config) is built")
    table = build_table(sde_x, sde_adj, predictor, corrector, snr, scale_eps, eps, probability_flow)
    sdesc = make_sampler_desc(sde_x, sde_adj, predictor, corrector, snr, scale_eps, eps, 1, probability_flow, denoise)
    steps = table.shape[0]
    N = int(shape_adj[-1])

    def sampler(model_x, model_adj, init_flags):
        dev = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
        eng = _engine_for(model_x, model_adj, N, dev)
        flags = init_flags.to(dev, torch.float32)
        seed = torch.initial_seed()
        iters = steps if max_iters is None else min(steps, max_iters)
        if shard is None:
            x, adj, xm, am = eng.sample(sdesc, table, flags, n_iters=iters, seed=seed, use_graph=use_graph)
        else:
            lo, hi = shard.bounds(flags.shape[0])
            x, adj, xm, am = eng.sample(sdesc, table, flags[lo:hi], n_iters=iters, seed=seed, graph_offset=lo,
                                        global_B=flags.shape[0], comm=shard.comm, use_graph=use_graph)
            # the final gather (every rank returns the whole batch, like the single-process reference): only the pair
            # that is returned travels
            if denoise:
                xm, am = [shard.gather(t, flags.shape[0]) for t in (xm, am)]
            else:
                x, adj = [shard.gather(t, flags.shape[0]) for t in (x, adj)]
        n_evals = 0 if s4 else steps * (int(n_steps) + 1)            # solver.py:195, :279
        return (xm if denoise else x), (am if denoise else adj), n_evals

    return sampler


def get_pc_sampler(sde_x, sde_adj, shape_x, shape_adj, predictor='Euler', corrector='None', snr=0.1, scale_eps=1.0,
                   n_steps=1, probability_flow=False, continuous=False, denoise=True, eps=1e-3, device='cuda', **kw):
    """solver.py:153-156."""
    return _make(sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps, probability_flow,
                 continuous, denoise, eps, device, False, **kw)


def S4_solver(sde_x, sde_adj, shape_x, shape_adj, predictor='None', corrector='None', snr=0.1, scale_eps=1.0,
              n_steps=1, probability_flow=False, continuous=False, denoise=True, eps=1e-3, device='cuda', **kw):
    """solver.py:200-203."""
    return _make(sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps, probability_flow,
                 continuous, denoise, eps, device, True, **kw)


def load_sampling_fn(config_train, config_module, config_sample, device):
    """utils/loader.py:121-147 (config objects with attribute or key access)."""
    from .sde import load_sde

    def g(o, k):
        return o[k] if isinstance(o, dict) else getattr(o, k)

    sde_x, sde_adj = load_sde(g(g(config_train, "sde"), "x")), load_sde(g(g(config_train, "sde"), "adj"))
    data = g(config_train, "data")
    n = g(data, "max_node_num")
    dev = f"cuda:{device[0]}" if isinstance(device, list) else device
    get = S4_solver if g(config_module, "predictor") == "S4" else get_pc_sampler
    bs = 10000 if g(data, "data") in ("QM9", "ZINC250k") else g(data, "batch_size")
    return get(sde_x=sde_x, sde_adj=sde_adj, shape_x=(bs, n, g(data, "max_feat_num")), shape_adj=(bs, n, n),
               predictor=g(config_module, "predictor"), corrector=g(config_module, "corrector"),
               snr=g(config_module, "snr"), scale_eps=g(config_module, "scale_eps"), n_steps=g(config_module, "n_steps"),
               probability_flow=g(config_sample, "probability_flow"), continuous=True,
               denoise=g(config_sample, "noise_removal"), eps=g(config_sample, "eps"), device=dev)
