"""Generate tests/golden/ from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # needs /root/reference

Produces
  tests/golden/ckpt/<name>.pt     the 7 shipped checkpoints re-saved as plain
                                  dict / fp32 tensors (no EasyDict pickle), so the
                                  GPU box -- which has no /root/reference -- can load them
  tests/golden/kat_<name>.npz     raw ScoreNetworkX*/ScoreNetworkA outputs of the reference
                                  modules on the deterministic SURVEY section-8c input recipe
  tests/golden/traj_<case>.npz    final (x, adj) of the reference's get_pc_sampler / S4_solver
                                  closures, with every torch.randn / randn_like call served by
                                  numpy's frozen RandomState(seed) stream so the same raw draws
                                  can be replayed through the oracle and injected into the
                                  CUDA path
  tests/golden/decode_zinc_v2.npz sampler.py:131-138 (quantize_mol, bond-class remap, one-hot, atom threshold + virtual
                                  atom) applied to the final state of the full ZINC250k v2 reference run
  tests/golden/c1_community_reference.npz  BASELINE config C1 run with the reference on the CPU (flags, thresholded
                                  samples, training graphs) for the statistical parity test
  tests/golden/c1_mmd_reference.npz  the reference's own evaluation/stats.py (degree / cluster / orbit / spectral MMD, ORCA orbit
                                  counts) on that run (mode `mmd`)
  tests/golden/nspdk_reference.npz  the reference's own NSPDK MMD (evaluation/mmd.py compute_nspdk_mmd over evaluation/eden.py) on the
                                  molecule graphs of the C2 / C3 reference runs (mode `nspdk`)
  tests/golden/dsm_<ckpt>_<sum|mean>.npz  loss_fn of losses.py:37-81 on a synthetic batch with recorded t / noise draws
  tests/golden/MANIFEST.json      what was generated, from which torch version

The reference code is imported and called, never copied.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import gdss_oracle as O          # noqa: E402
from oracle import refimport                 # noqa: E402
from oracle.cases import CKPTS, SHORT_CASES, FULL_CASES, NumpyDraws, plain   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def convert_ckpts():
    os.makedirs(os.path.join(GOLD, "ckpt"), exist_ok=True)
    out = {}
    for name, rel in CKPTS.items():
        ck = refimport.load_ckpt(rel)
        use_ema = bool(ck["model_config"].get("sample", {}).get("use_ema", False))
        d = {
            "model_config": plain(ck["model_config"]),
            "params_x": plain(ck["params_x"]),
            "params_adj": plain(ck["params_adj"]),
            "x_state_dict": {k: v.detach().float().clone() for k, v in ck["x_state_dict"].items()},
            "adj_state_dict": {k: v.detach().float().clone() for k, v in ck["adj_state_dict"].items()},
        }
        if use_ema:     # only ENZYMES / grid sample with EMA weights (sampler.py:47-52)
            for key in ("ema_x", "ema_adj"):
                e = ck[key]
                d[key] = {"decay": float(e["decay"]), "num_updates": int(e["num_updates"]),
                          "shadow_params": [p.detach().float().clone() for p in e["shadow_params"]]}
        path = os.path.join(GOLD, "ckpt", f"{name}.pt")
        torch.save(d, path)
        out[name] = {"source": f"checkpoints/{rel}.pth", "bytes": os.path.getsize(path), "ema": use_ema}
    return out


def ref_models(name, ema=False):
    from utils.loader import load_model
    ck = refimport.load_ckpt(CKPTS[name])
    px, pa = dict(ck["params_x"]), dict(ck["params_adj"])
    mx, ma = load_model(px), load_model(pa)
    mx.load_state_dict(ck["x_state_dict"])
    ma.load_state_dict(ck["adj_state_dict"])
    if ema:
        from utils.ema import ExponentialMovingAverage
        for m, key in ((mx, "ema_x"), (ma, "ema_adj")):
            e = ExponentialMovingAverage(m.parameters(), decay=0.999)
            e.load_state_dict(ck[key])
            e.copy_to(m.parameters())
    return mx.eval(), ma.eval(), px, pa


def make_kat():
    out = {}
    for name in CKPTS:
        mx, ma, px, pa = ref_models(name)
        N, F = pa["max_node_num"], pa["max_feat_num"]
        B = 4 if N < 200 else 2
        x, adj, flags = O.kat_inputs(N, F, B)
        with torch.no_grad():
            sx, sa = mx(x, adj, flags), ma(x, adj, flags)
        np.savez_compressed(os.path.join(GOLD, f"kat_{name}.npz"), score_x=sx.numpy(), score_adj=sa.numpy(),
                            B=B, N=N, F=F)
        out[name] = {"B": B, "abs_sum_x": float(sx.double().abs().sum()), "abs_sum_adj": float(sa.double().abs().sum())}
        print("kat", name, out[name], flush=True)
    return out


@contextlib.contextmanager
def numpy_randn(seed):
    """Serve torch.randn / torch.randn_like from RandomState(seed) while the reference runs."""
    draws = NumpyDraws(seed)
    orig_randn, orig_like = torch.randn, torch.randn_like

    def randn(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        return draws(shape)

    def randn_like(t, **kw):
        return draws(tuple(t.shape))

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield draws
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like


def run_reference_case(case):
    import sde as RSDE
    import solver as RS
    mx, ma, px, pa = ref_models(case["ckpt"], ema=case.get("ema", False))
    N, F, B = pa["max_node_num"], pa["max_feat_num"], case["B"]
    _, _, flags = O.kat_inputs(N, F, B)
    cls = {"VP": RSDE.VPSDE, "VE": RSDE.VESDE, "subVP": RSDE.subVPSDE}
    sx = cls[case["sde_x"][0]](case["sde_x"][1], case["sde_x"][2], case["scales"])
    sa = cls[case["sde_adj"][0]](case["sde_adj"][1], case["sde_adj"][2], case["scales"])
    get = RS.S4_solver if case["predictor"] == "S4" else RS.get_pc_sampler
    fn = get(sx, sa, (B, N, F), (B, N, N), predictor=case["predictor"], corrector=case["corrector"],
             snr=case["snr"], scale_eps=case["scale_eps"], n_steps=case.get("n_steps", 1),
             probability_flow=case.get("probability_flow", False), continuous=True,
             denoise=case.get("denoise", True), eps=1e-4, device="cpu")
    with numpy_randn(case["seed"]), contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()):
        x, adj, n_evals = fn(mx, ma, flags)
    return flags, x, adj, n_evals


def make_traj(missing_only=False):
    out = {}
    for case in SHORT_CASES + FULL_CASES:
        if missing_only and os.path.exists(os.path.join(GOLD, f"traj_{case['name']}.npz")):
            continue
        flags, x, adj, n_evals = run_reference_case(case)
        assert case["ckpt"] == "enzymes" or (torch.isfinite(x).all() and torch.isfinite(adj).all())
        np.savez_compressed(os.path.join(GOLD, f"traj_{case['name']}.npz"), flags=flags.numpy(), x=x.numpy(),
                            adj=adj.numpy(), n_evals=n_evals, case=json.dumps(case))
        out[case["name"]] = {"n_evals": int(n_evals), "adj_gt_half": float((adj > 0.5).float().mean())}
        print("traj", case["name"], out[case["name"]], flush=True)
    return out


class _EveryNth(list):
    """trace sink that keeps the state after every n-th solver step only."""

    def __init__(self, n):
        super().__init__()
        self.n, self.count = n, 0

    def append(self, rec):
        self.count += 1
        if self.count % self.n == 0:
            super().append({k: rec[k].clone() for k in ("x", "adj", "x_mean", "adj_mean")})


SEGMENT_CASES = ["zinc_v2_rev_lang_full_b32"]


def make_segments_s4_ref(name="enzymes_s4_full_b4", every=100):
    """The same for an S4 run whose oracle replay is close but not bit-identical (ENZYMES, N = 125): the intermediate states are
    taken from the UNMODIFIED reference run itself -- S4_solver evaluates model_x exactly once per step on the current state
    (solver.py:228), so a forward pre-hook on model_x sees the state after k steps at its k-th call."""
    import sde as RSDE
    import solver as RS
    from oracle.cases import ALL_CASES
    case = ALL_CASES[name]
    z = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
    mx, ma, px, pa = ref_models(case["ckpt"], ema=case.get("ema", False))
    N, F, B = pa["max_node_num"], pa["max_feat_num"], case["B"]
    flags = torch.from_numpy(z["flags"])
    cls = {"VP": RSDE.VPSDE, "VE": RSDE.VESDE, "subVP": RSDE.subVPSDE}
    sx = cls[case["sde_x"][0]](case["sde_x"][1], case["sde_x"][2], case["scales"])
    sa = cls[case["sde_adj"][0]](case["sde_adj"][1], case["sde_adj"][2], case["scales"])
    fn = RS.S4_solver(sx, sa, (B, N, F), (B, N, N), predictor="S4", corrector=case["corrector"], snr=case["snr"],
                      scale_eps=case["scale_eps"], n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-4,
                      device="cpu")
    xs, as_, calls = [], [], [0]

    def hook(mod, args):
        if calls[0] > 0 and calls[0] % every == 0:
            xs.append(args[0].detach().clone())
            as_.append(args[1].detach().clone())
        calls[0] += 1

    h = mx.register_forward_pre_hook(hook)
    with numpy_randn(case["seed"]), contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        x, adj, _ = fn(mx, ma, flags)
    h.remove()
    same = lambda a, b: torch.equal(a.contiguous().view(torch.int32), torch.from_numpy(b).contiguous().view(torch.int32))
    assert same(x, z["x"]) and same(adj, z["adj"]) and calls[0] == case["scales"], "the re-run does not reproduce the fixture"
    # the state after the last step is not seen by a hook: segment k < last ends at xs[k]; the last one at the fixture's means
    np.savez_compressed(os.path.join(GOLD, f"seg_{name}.npz"), every=every, x=torch.stack(xs).numpy(), adj=torch.stack(as_).numpy(),
                        x_mean=z["x"], adj_mean=z["adj"])
    print("segments (reference hooks)", name, len(xs), flush=True)
    return {name: {"segments": len(xs) + 1, "every": every, "source": "reference forward pre-hook"}}


def make_segments(every=50):
    """Intermediate states of the full-length runs, every `every` solver steps, for segment-wise (teacher-forced) parity of
    trajectories that are chaotic over 1000 steps (a 2e-7 relative perturbation of the prior changes the reference's own
    final ZINC250k sample by O(1): tools/chaos_probe.py).  Produced by the oracle, whose final state must equal the
    UNMODIFIED reference's fixture bit for bit first (same torch ops in the same order), which pins every intermediate."""
    from oracle.cases import ALL_CASES, case_weights
    out = {}
    for name in SEGMENT_CASES:
        case = ALL_CASES[name]
        z = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
        flags = torch.from_numpy(z["flags"])
        sdx, px, sda, pa = case_weights(case)
        N, F, B = pa["max_node_num"], pa["max_feat_num"], case["B"]
        sx, sa = O.SDESpec(*case["sde_x"], case["scales"]), O.SDESpec(*case["sde_adj"], case["scales"])
        sink = _EveryNth(every)
        draw = NumpyDraws(case["seed"])
        with torch.no_grad():
            if case["predictor"] == "S4":
                x, adj, _ = O.s4_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw, snr=case["snr"],
                                        scale_eps=case["scale_eps"], denoise=True, eps=1e-4, trace=sink)
            else:
                x, adj, _ = O.pc_sample(sx, sa, sdx, px, sda, pa, (B, N, F), (B, N, N), flags, draw, predictor=case["predictor"],
                                        corrector=case["corrector"], snr=case["snr"], scale_eps=case["scale_eps"], denoise=True,
                                        eps=1e-4, trace=sink)
        same = lambda a, b: torch.equal(a.contiguous().view(torch.int32), torch.from_numpy(b).contiguous().view(torch.int32))
        assert same(x, z["x"]) and same(adj, z["adj"]), f"{name}: the oracle is not bit-identical to the reference run"
        np.savez_compressed(os.path.join(GOLD, f"seg_{name}.npz"), every=every,
                            x=torch.stack([r["x"] for r in sink]).numpy(), adj=torch.stack([r["adj"] for r in sink]).numpy(),
                            x_mean=sink[-1]["x_mean"].numpy(), adj_mean=sink[-1]["adj_mean"].numpy())
        out[name] = {"segments": len(sink), "every": every}
        print("segments", name, out[name], flush=True)
    return out


LAYER_KAT = ["qm9", "zinc250k_v2", "community_small"]


def make_layer_kat():
    """Op-level known answers (SURVEY 8c protocol (i)): outputs of the reference's own sub-modules on the KAT inputs, captured
    with forward hooks -- every AttentionLayer of ScoreNetworkA (node output, adjacency-channel output), the Q / K / V
    DenseGCNConv outputs of its LAST layer, pow_tensor's stack, and every layer's node features of the X network (as they
    enter the final concatenation)."""
    from utils.graph_utils import pow_tensor
    out = {}
    for name in LAYER_KAT:
        mx, ma, px, pa = ref_models(name)
        N, F, B = pa["max_node_num"], pa["max_feat_num"], 2
        x, adj, flags = O.kat_inputs(N, F, B)
        rec = {}
        hooks = []
        for l, layer in enumerate(ma.layers):
            hooks.append(layer.register_forward_hook(lambda m, i, o, l=l: rec.update({f"a_x_{l}": o[0].detach().numpy(), f"a_adj_{l}": o[1].detach().numpy()})))
        last = ma.layers[-1]
        for c, att in enumerate(last.attn):
            for nm in "qkv":
                hooks.append(getattr(att, f"gnn_{nm}").register_forward_hook(
                    lambda m, i, o, c=c, nm=nm: rec.update({f"a_last_{nm}_{c}": o.detach().numpy()})))
        with torch.no_grad():
            sa = ma(x, adj, flags)
        for h in hooks:
            h.remove()
        # X network: the tensors that enter its final concatenation (ScoreNetwork_X.py:40, :84)
        xs, hooks = [], []
        orig_cat = torch.cat

        def spy(tensors, dim=0, **kw):
            if dim == -1 and len(tensors) == len(mx.layers) + 1 and tensors[0].shape == x.shape:
                xs.extend(t.detach().numpy() for t in tensors)
            return orig_cat(tensors, dim=dim, **kw)

        torch.cat = spy
        try:
            with torch.no_grad():
                sxo = mx(x, adj, flags)
        finally:
            torch.cat = orig_cat
        assert len(xs) == len(mx.layers) + 1, "the X network's concatenation was not seen"
        for l, t in enumerate(xs[1:]):
            rec[f"x_x_{l}"] = t
        rec["pow"] = pow_tensor(adj, pa["c_init"]).numpy()
        rec["score_x"], rec["score_adj"] = sxo.numpy(), sa.numpy()
        np.savez_compressed(os.path.join(GOLD, f"kat_layers_{name}.npz"), B=B, N=N, F=F, a_layers=len(ma.layers), x_layers=len(mx.layers),
                            **rec)
        out[name] = sorted(rec.keys())
        print("layer kat", name, len(rec), flush=True)
    return out


def make_decode():
    """sampler.py:131-138 executed literally on the final state of the full ZINC250k v2 reference run (quantize_mol is
    the reference's own function; the remaining lines are inline in Sampler_mol.sample, which needs rdkit to import)."""
    from utils.graph_utils import quantize_mol
    z = np.load(os.path.join(GOLD, "traj_zinc_v2_rev_lang_full.npz"))
    x, adj = torch.from_numpy(z["x"]), torch.from_numpy(z["adj"])
    samples_int = quantize_mol(adj.clone())
    samples_int = samples_int - 1
    samples_int[samples_int == -1] = 3
    a = torch.nn.functional.one_hot(torch.tensor(samples_int), num_classes=4).permute(0, 3, 1, 2)
    xi = torch.where(x > 0.5, 1, 0)
    xi = torch.concat([xi, 1 - xi.sum(dim=-1, keepdim=True)], dim=-1)
    np.savez_compressed(os.path.join(GOLD, "decode_zinc_v2.npz"), x=xi.numpy().astype(np.int8), adj=a.numpy().astype(np.int8))
    return {"decode_zinc_v2": [list(xi.shape), list(a.shape)]}


def main():
    refimport.activate()
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"torch": torch.__version__, "numpy": np.__version__, "reference": refimport.REF}
    manifest["ckpt"] = convert_ckpts()
    manifest["kat"] = make_kat()
    manifest["traj"] = make_traj()
    manifest["decode"] = make_decode()
    manifest["c1"] = make_c1_stats()
    manifest["dsm"] = make_dsm()
    manifest["mol_stats"] = make_mol_stats()
    manifest["dsm_grad"] = make_dsm_grads()
    manifest["train_steps"] = make_train_steps()
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


def make_c1_stats():
    """BASELINE config C1 exactly as SURVEY 8d specifies it, run with the UNMODIFIED reference on the CPU: community_small,
    B = 100, Euler + Langevin (snr .05, scale_eps .7), 1000 steps, flags = node counts of training graphs
    data/community_small.pkl[20:] drawn with RandomState(11).randint, torch.manual_seed(11).  Stored: the flags, the
    thresholded adjacency of the samples and of the training graphs (for degree / edge statistics in the GPU tests)."""
    import pickle
    import sde as RSDE
    import solver as RS
    with open(os.path.join(refimport.REF, "data", "community_small.pkl"), "rb") as f:
        graphs = pickle.load(f)
    train = graphs[int(0.2 * len(graphs)):]
    N, F, B = 20, 10, 100
    idx = np.random.RandomState(11).randint(0, len(train), B)
    counts = np.array([train[i].number_of_nodes() for i in idx])
    flags = (torch.arange(N)[None, :] < torch.from_numpy(counts)[:, None]).float()
    import networkx as nx
    train_adj = np.zeros((len(train), N, N), dtype=np.int8)
    for k, gph in enumerate(train):
        a = nx.to_numpy_array(gph)
        train_adj[k, :a.shape[0], :a.shape[1]] = a > 0
    mx, ma, px, pa = ref_models("community_small")
    sx, sa = RSDE.VPSDE(.1, 1., 1000), RSDE.VPSDE(.1, 1., 1000)
    fn = RS.get_pc_sampler(sx, sa, (B, N, F), (B, N, N), predictor="Euler", corrector="Langevin", snr=.05, scale_eps=.7,
                           n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-4, device="cpu")
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        x, adj, _ = fn(mx, ma, flags)
    q = (adj >= 0.5).numpy().astype(np.int8)
    np.savez_compressed(os.path.join(GOLD, "c1_community_reference.npz"), counts=counts.astype(np.int16), samples=q,
                        train=train_adj, train_counts=np.array([g.number_of_nodes() for g in train], dtype=np.int16))
    e = q.sum(axis=(1, 2)) / 2
    return {"mean_nodes": float(counts.mean()), "edges_mean": float(e.mean()), "edges_std": float(e.std())}


def make_mol_stats():
    """BASELINE configs C2 / C3 at a reduced batch, run with the UNMODIFIED reference on the CPU for the full 1000 steps with
    torch's own generator (torch.manual_seed(42)) and the SURVEY 8d node-count recipes (RandomState(42)):
      C2  QM9, S4_solver (snr .2, scale_eps .7, VE / VE), B = 384
      C3  ZINC250k v2, get_pc_sampler Reverse + Langevin (snr .2, scale_eps .8, VP / VE), B = 24
    Stored: node counts, quantize_mol(adj) (bond orders 0..3) and x > 0.5 (atom channels) of the samples -- the statistics the
    GPU test compares (the CUDA sampler draws from Philox, so only distributions can agree)."""
    import sde as RSDE
    import solver as RS
    from utils.graph_utils import quantize_mol
    out = {}
    rs = np.random.RandomState(42)
    specs = [("qm9", "qm9", 9, 4, 384, rs.choice([9, 8, 7, 6, 5], size=384, p=[.83, .14, .025, .004, .001]),
              (RSDE.VESDE(.1, 1., 1000), RSDE.VESDE(.1, 1., 1000)), ("S4", "None", .2, .7)),
             ("zinc250k_v2", "zinc_v2", 38, 9, 24,
              np.clip(np.round(np.random.RandomState(42).normal(23.2, 4.5, 24)), 6, 38).astype(int),
              (RSDE.VPSDE(.1, 1., 1000), RSDE.VESDE(.2, 1., 1000)), ("Reverse", "Langevin", .2, .8))]
    for ckpt, tag, N, F, B, counts, (sx, sa), (pred, corr, snr, seps) in specs:
        flags = (torch.arange(N)[None, :] < torch.from_numpy(np.asarray(counts))[:, None]).float()
        mx, ma, px, pa = ref_models(ckpt)
        get = RS.S4_solver if pred == "S4" else RS.get_pc_sampler
        fn = get(sx, sa, (B, N, F), (B, N, N), predictor=pred, corrector=corr, snr=snr, scale_eps=seps, n_steps=1,
                 probability_flow=False, continuous=True, denoise=True, eps=1e-4, device="cpu")
        torch.manual_seed(42)
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            x, adj, _ = fn(mx, ma, flags)
        assert torch.isfinite(x).all() and torch.isfinite(adj).all()
        bonds = quantize_mol(adj.clone()).astype(np.int8)
        atoms = (x > 0.5).numpy().astype(np.int8)
        np.savez_compressed(os.path.join(GOLD, f"stats_{tag}_reference.npz"), counts=np.asarray(counts, dtype=np.int16),
                            bonds=bonds, atoms=atoms)
        out[tag] = {"B": B, "bond_hist": np.bincount(bonds.reshape(-1), minlength=4).tolist(),
                    "atom_on": int(atoms.sum())}
        print("mol stats", tag, out[tag], flush=True)
    return out


DSM_CASES = [("qm9", ("VE", .1, 1.), ("VE", .1, 1.)), ("zinc250k_v2", ("VP", .1, 1.), ("VE", .2, 1.)),
             ("community_small", ("VP", .1, 1.), ("VP", .1, 1.)), ("ego_small", ("subVP", .1, 1.), ("VE", .1, 1.))]


def dsm_inputs(N, F, B, seed):
    """Synthetic training batch of the reference's shape: one-hot node features, symmetric integer bond orders, ragged
    node counts (real data is not shipped)."""
    rs = np.random.RandomState(seed)
    x = np.zeros((B, N, F), dtype=np.float32)
    adj = np.zeros((B, N, N), dtype=np.float32)
    for b in range(B):
        n = rs.randint(max(3, N // 2), N + 1)
        x[b, np.arange(n), rs.randint(0, F, n)] = 1
        a = np.triu((rs.rand(n, n) < 0.3) * rs.randint(1, 4, (n, n)), 1).astype(np.float32)
        adj[b, :n, :n] = a + a.T
        for i in range(n):                               # every node keeps at least one bond (node_flags sees it)
            if adj[b, i, :n].sum() == 0:
                j = (i + 1) % n
                adj[b, i, j] = adj[b, j, i] = 1
    return torch.from_numpy(x), torch.from_numpy(adj)


def make_dsm():
    """loss_fn of the UNMODIFIED reference (losses.py:37-81) with torch.rand / torch.randn_like served from a recorded
    numpy stream; stores the inputs, the three draws and the two loss values."""
    import losses as RL
    import sde as RSDE
    out = {}
    cls = {"VP": RSDE.VPSDE, "VE": RSDE.VESDE, "subVP": RSDE.subVPSDE}
    for k, (name, sx_, sa_) in enumerate(DSM_CASES):
        mx, ma, px, pa = ref_models(name)
        N, F, B = pa["max_node_num"], pa["max_feat_num"], 6
        x, adj = dsm_inputs(N, F, B, 300 + k)
        sx, sa = cls[sx_[0]](sx_[1], sx_[2], 1000), cls[sa_[0]](sa_[1], sa_[2], 1000)
        for rm in (False, True):
            fn = RL.get_sde_loss_fn(sx, sa, train=False, reduce_mean=rm, continuous=True, likelihood_weighting=False, eps=1e-5)
            rs = np.random.RandomState(400 + k)
            tape = []
            orig_rand, orig_like = torch.rand, torch.randn_like

            def rand(*shape, **kw):
                v = torch.from_numpy(rs.rand(*shape).astype(np.float32)); tape.append(v); return v

            def randn_like(tt, **kw):
                v = torch.from_numpy(rs.randn(*tt.shape).astype(np.float32)); tape.append(v); return v

            torch.rand, torch.randn_like = rand, randn_like
            try:
                with torch.no_grad():
                    lx, la = fn(mx, ma, x, adj)
            finally:
                torch.rand, torch.randn_like = orig_rand, orig_like
            assert len(tape) == 3
            key = f"{name}_{'mean' if rm else 'sum'}"
            np.savez_compressed(os.path.join(GOLD, f"dsm_{key}.npz"), x=x.numpy(), adj=adj.numpy(), t=tape[0].numpy(),
                                raw_x=tape[1].numpy(), raw_adj=tape[2].numpy(), loss_x=float(lx), loss_adj=float(la),
                                sde=json.dumps([sx_, sa_]), reduce_mean=rm)
            out[key] = [float(lx), float(la)]
            print("dsm", key, out[key], flush=True)
    return out


def make_dsm_grads():
    """Gradients of the two DSM losses (losses.py:37-81, train = True, reduce_mean = False) with respect to every parameter
    of both score networks, from the UNMODIFIED reference's autograd, for the inputs / recorded draws of the existing
    dsm_*_sum.npz fixtures (same seeds).  Stored flat in parameters() order: the parity target of the backward kernels
    (trainer.py:57-65: loss_x.backward(); loss_adj.backward())."""
    import losses as RL
    import sde as RSDE
    out = {}
    cls = {"VP": RSDE.VPSDE, "VE": RSDE.VESDE, "subVP": RSDE.subVPSDE}
    for k, (name, sx_, sa_) in enumerate(DSM_CASES):
        mx, ma, px, pa = ref_models(name)
        z = np.load(os.path.join(GOLD, f"dsm_{name}_sum.npz"))
        x, adj = torch.from_numpy(z["x"]), torch.from_numpy(z["adj"])
        sx, sa = cls[sx_[0]](sx_[1], sx_[2], 1000), cls[sa_[0]](sa_[1], sa_[2], 1000)
        fn = RL.get_sde_loss_fn(sx, sa, train=True, reduce_mean=False, continuous=True, likelihood_weighting=False, eps=1e-5)
        tape = [torch.from_numpy(z["t"]), torch.from_numpy(z["raw_x"]), torch.from_numpy(z["raw_adj"])]
        it = iter(tape)
        orig_rand, orig_like = torch.rand, torch.randn_like
        torch.rand, torch.randn_like = (lambda *a, **kw: next(it)), (lambda tt, **kw: next(it))
        try:
            for m in (mx, ma):
                for p_ in m.parameters():
                    p_.requires_grad_(True)
                m.zero_grad()
            lx, la = fn(mx, ma, x, adj)
            lx.backward()
            la.backward()
        finally:
            torch.rand, torch.randn_like = orig_rand, orig_like
        assert abs(float(lx) - float(z["loss_x"])) <= 1e-5 * abs(float(z["loss_x"])) + 1e-6
        # parameters the loss never reaches keep grad = None in torch (the last AttentionLayer's multi_channel MLP of
        # ScoreNetworkA: its x output is dropped, ScoreNetwork_A.py:139-141) and torch.optim.Adam skips them: stored as
        # zeros with a per-element mask
        def flat(m):
            g = [(p_.grad if p_.grad is not None else torch.zeros_like(p_)).reshape(-1) for p_ in m.parameters()]
            has = [torch.full((p_.numel(),), 0 if p_.grad is None else 1, dtype=torch.uint8) for p_ in m.parameters()]
            return torch.cat(g).numpy(), torch.cat(has).numpy()
        (gx, hx), (ga, ha) = flat(mx), flat(ma)
        np.savez_compressed(os.path.join(GOLD, f"dsm_grad_{name}_sum.npz"), grad_x=gx, grad_adj=ga, has_grad_x=hx, has_grad_adj=ha,
                            loss_x=float(lx), loss_adj=float(la))
        out[name] = [int(gx.size), int(ga.size), float(np.linalg.norm(gx)), float(np.linalg.norm(ga))]
        print("dsm grad", name, out[name], flush=True)
        mx.eval(); ma.eval()
    return out


def make_train_steps():
    """Two consecutive iterations of the UNMODIFIED reference's training loop body (trainer.py:57-79) on ZINC250k v2 with the
    shipped config's optimizer settings (config/zinc250k.yaml:35-47: Adam lr 0.005, weight_decay 1e-4, grad_norm 1.0, ema
    0.999; optimizers as utils/loader.py:51-64 builds them, EMA = utils/ema.py), starting from the checkpoint, on the batch /
    recorded draws of dsm_zinc250k_v2_sum.npz (the same draws in both iterations).  Stored: losses, pre-clip gradient norms,
    flat parameters and EMA shadows of both networks after each iteration."""
    import losses as RL
    import sde as RSDE
    from utils.ema import ExponentialMovingAverage
    name = "zinc250k_v2"
    mx, ma, px, pa = ref_models(name)
    z = np.load(os.path.join(GOLD, f"dsm_{name}_sum.npz"))
    x, adj = torch.from_numpy(z["x"]), torch.from_numpy(z["adj"])
    sx, sa = RSDE.VPSDE(.1, 1., 1000), RSDE.VESDE(.2, 1., 1000)
    fn = RL.get_sde_loss_fn(sx, sa, train=True, reduce_mean=False, continuous=True, likelihood_weighting=False, eps=1e-5)
    for m in (mx, ma):
        for p_ in m.parameters():
            p_.requires_grad_(True)
    ox = torch.optim.Adam(mx.parameters(), lr=0.005, weight_decay=0.0001)
    oa = torch.optim.Adam(ma.parameters(), lr=0.005, weight_decay=0.0001)
    ex, ea = ExponentialMovingAverage(mx.parameters(), decay=0.999), ExponentialMovingAverage(ma.parameters(), decay=0.999)
    rec = {}
    orig_rand, orig_like = torch.rand, torch.randn_like
    try:
        for it_ in range(2):
            tape = iter([torch.from_numpy(z["t"]), torch.from_numpy(z["raw_x"]), torch.from_numpy(z["raw_adj"])])
            torch.rand, torch.randn_like = (lambda *a, **kw: next(tape)), (lambda tt, **kw: next(tape))
            mx.train(); ma.train()
            ox.zero_grad(); oa.zero_grad()
            lx, la = fn(mx, ma, x, adj)
            lx.backward(); la.backward()
            nx_ = torch.nn.utils.clip_grad_norm_(mx.parameters(), 1.0)
            na_ = torch.nn.utils.clip_grad_norm_(ma.parameters(), 1.0)
            ox.step(); oa.step()
            ex.update(mx.parameters()); ea.update(ma.parameters())
            flat = lambda ps: torch.cat([p_.detach().reshape(-1) for p_ in ps]).numpy().copy()
            rec[f"loss_{it_}"] = np.array([float(lx), float(la)])
            rec[f"gnorm_{it_}"] = np.array([float(nx_), float(na_)])
            rec[f"px_{it_}"], rec[f"pa_{it_}"] = flat(mx.parameters()), flat(ma.parameters())
            rec[f"ex_{it_}"], rec[f"ea_{it_}"] = flat(ex.shadow_params), flat(ea.shadow_params)
    finally:
        torch.rand, torch.randn_like = orig_rand, orig_like
    np.savez_compressed(os.path.join(GOLD, f"train_steps_{name}.npz"), **rec)
    out = {k: rec[k].tolist() for k in rec if k.startswith(("loss", "gnorm"))}
    print("train steps", out, flush=True)
    return out


def make_mmd_reference():
    """The reference's own evaluation (evaluation/stats.py `degree_stats`, `clustering_stats`, `orbit_stats_all`, `spectral_stats`
    with the kernels of utils/loader.py:199-206) run UNMODIFIED on the C1 reference run: generated samples (through the
    reference's `adjs_to_graphs` convention: isolated nodes removed) against the training graphs.  What has to be supplied from
    outside: `np.float` (removed from numpy), `nx.from_numpy_matrix` (removed from networkx 3; `from_numpy_array` is the same
    function), `pyemd.emd` (not installed: the exact 1-D earth mover's distance for the Toeplitz ground distance |i - j| / s,
    checked here against the transportation LP solved by scipy on a sample of the actual histogram pairs) and the ORCA binary
    (built from the reference's own orca.cpp by oracle/Makefile into oracle/_ref).  Stored: the four MMD values and ORCA's
    per-node orbit counts of every graph, which pin oracle/stats.py."""
    import subprocess
    import networkx as nx
    import pyemd
    from scipy.optimize import linprog
    np.float = float
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)

    def emd_exact_1d(x, y, dist):
        return float(np.abs(np.cumsum(np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64))).sum() * dist[0, 1])

    def emd_lp(x, y, dist):
        n = len(x)
        a_eq = np.zeros((2 * n, n * n))
        for i in range(n):
            a_eq[i, i * n:(i + 1) * n] = 1
            a_eq[n + i, i::n] = 1
        r = linprog(dist.reshape(-1), A_eq=a_eq, b_eq=np.concatenate([x, y]), bounds=(0, None), method="highs")
        return float(r.fun)

    emd = emd_exact_1d          # (the reference evaluates the kernels in worker processes: evaluation/mmd.py:84-87)
    pyemd.emd = emd
    import evaluation.stats as S
    import evaluation.mmd as M
    S.ORCA_DIR = os.path.join(HERE, "_ref")
    z = np.load(os.path.join(GOLD, "c1_community_reference.npz"))
    train = [nx.from_numpy_array(z["train"][k][:n, :n].astype(np.float64)) for k, n in enumerate(z["train_counts"])]

    def adjs_to_graphs(adjs):               # utils/graph_utils.py:165-176 with from_numpy_array for the removed from_numpy_matrix
        out = []
        for adj in adjs:
            g = nx.from_numpy_array(adj)
            g.remove_edges_from(nx.selfloop_edges(g))
            g.remove_nodes_from(list(nx.isolates(g)))
            if g.number_of_nodes() < 1:
                g.add_node(1)
            out.append(g)
        return out

    pred = adjs_to_graphs(z["samples"].astype(np.float64))
    with contextlib.redirect_stdout(io.StringIO()):
        res = {"degree": S.degree_stats(train, pred, M.gaussian_emd, is_parallel=False),
               "cluster": S.clustering_stats(train, pred, M.gaussian_emd, is_parallel=False),
               "orbit": S.orbit_stats_all(train, pred, M.gaussian),
               "spectral": S.spectral_stats(train, pred, M.gaussian_emd, is_parallel=False)}
    res = {k: round(float(v), 6) for k, v in res.items()}
    # the closed-form EMD against the transportation LP on 40 of the actual degree-histogram pairs
    from scipy.linalg import toeplitz
    rs = np.random.RandomState(0)
    lp_checks = []
    for _ in range(40):
        hx, hy = S.degree_worker(train[rs.randint(len(train))]), S.degree_worker(pred[rs.randint(len(pred))])
        hx, hy = M.process_tensor(hx / hx.sum(), hy / hy.sum())
        dist = toeplitz(range(len(hx))).astype(np.float64)
        lp_checks.append(abs(emd_exact_1d(hx, hy, dist) - emd_lp(np.asarray(hx, dtype=np.float64), np.asarray(hy, dtype=np.float64), dist)))
    orb = [S.orca(g) for g in train + pred]
    sizes = np.array([o.shape[0] for o in orb], dtype=np.int32)
    np.savez_compressed(os.path.join(GOLD, "c1_mmd_reference.npz"), degree=res["degree"], cluster=res["cluster"], orbit=res["orbit"],
                        spectral=res["spectral"], orbit_counts=np.concatenate(orb).astype(np.int32), orbit_sizes=sizes,
                        n_train=len(train), emd_lp_max_diff=max(lp_checks), emd_lp_checks=len(lp_checks))
    res.update(emd_lp_checks=len(lp_checks), emd_lp_max_diff=float(max(lp_checks)))
    return res


def _mol_nx(graphs, symbols):
    """utils/mol_utils.py:161-183 `mols_to_nx` on (atomic numbers, bond orders): node label = the atom's symbol as sampler.py
    has it (symbols=True) or its atomic number (an integer label: Python's hash of it does not depend on PYTHONHASHSEED)."""
    import networkx as nx
    sym = {6: "C", 7: "N", 8: "O", 9: "F", 15: "P", 16: "S", 17: "Cl", 35: "Br", 53: "I"}
    out = []
    for lab, b in graphs:
        g = nx.Graph()
        for i, z in enumerate(lab):
            g.add_node(i, label=sym[int(z)] if symbols else int(z))
        for i in range(len(lab)):
            for j in range(i + 1, len(lab)):
                if b[i, j]:
                    g.add_edge(i, j, label=int(b[i, j]))
        out.append(g)
    return out


def _nspdk_halves(tag):
    from oracle import stats as OS
    z = np.load(os.path.join(GOLD, f"stats_{tag}_reference.npz"))
    gs = [g for g in OS.mol_graphs(z["atoms"], z["bonds"], OS.QM9_ATOMS if tag == "qm9" else OS.ZINC_ATOMS) if len(g[0]) > 0]
    return gs[:len(gs) // 2], gs[len(gs) // 2:]


def nspdk_symbol_run(tag):
    """one process = one PYTHONHASHSEED: the reference's NSPDK MMD with atom-symbol labels, as sampler.py:151 evaluates it."""
    import evaluation.mmd as M
    a, b = _nspdk_halves(tag)
    print(repr(float(M.compute_nspdk_mmd(_mol_nx(a, True), _mol_nx(b, True), metric="nspdk", is_hist=False, n_jobs=1))))


def make_nspdk_reference():
    """The reference's own NSPDK MMD (evaluation/mmd.py:115-127 `compute_nspdk_mmd` over evaluation/eden.py `vectorize`, run
    UNMODIFIED) on the molecule graphs of the C2 / C3 reference runs (tests/golden/stats_*_reference.npz, decoded as
    utils/mol_utils.py `construct_mol` reads them, first half of the run against the second half):
      * with integer node labels (atomic numbers): value + the feature vectors of the first graphs -> pins oracle/stats.py exactly;
      * with the atom-symbol labels sampler.py uses, in three processes with PYTHONHASHSEED 0 / 1 / 2: the label -> code map and
        with it the 16-bit feature collisions change from process to process; the spread is the reference's own noise floor."""
    import subprocess
    import evaluation.mmd as M
    from evaluation.eden import vectorize
    out, arrays = {}, {}
    for tag in ("qm9", "zinc_v2"):
        a, b = _nspdk_halves(tag)
        val = float(M.compute_nspdk_mmd(_mol_nx(a, False), _mol_nx(b, False), metric="nspdk", is_hist=False, n_jobs=1))
        feats = vectorize(_mol_nx(a[:6], False), complexity=4, discrete=True).tocsr()
        sym = []
        for seed in ("0", "1", "2"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "nspdk_symbols", tag], capture_output=True, text=True,
                               check=True, env=dict(os.environ, PYTHONHASHSEED=seed))
            sym.append(float(r.stdout.strip().splitlines()[-1]))
        arrays.update({f"{tag}_mmd_int": val, f"{tag}_mmd_symbols": np.array(sym), f"{tag}_n": np.array([len(a), len(b)]),
                       f"{tag}_feat_indptr": feats.indptr.astype(np.int32), f"{tag}_feat_indices": feats.indices.astype(np.int32),
                       f"{tag}_feat_data": feats.data})
        out[tag] = {"mmd_int_labels": val, "mmd_symbol_labels_hashseed_0_1_2": sym, "graphs": [len(a), len(b)]}
    np.savez_compressed(os.path.join(GOLD, "nspdk_reference.npz"), **arrays)
    return out


def only(which):
    """python oracle/make_golden.py decode | c1 | dsm | mol | dsmgrad | train : just one of the small fixtures."""
    refimport.activate()
    torch.set_num_threads(os.cpu_count())
    mpath = os.path.join(GOLD, "MANIFEST.json")
    manifest = json.load(open(mpath))
    if which == "decode":
        manifest["decode"] = make_decode()
    elif which == "dsm":
        manifest["dsm"] = make_dsm()
    elif which == "mol":
        manifest["mol_stats"] = make_mol_stats()
    elif which == "dsmgrad":
        manifest["dsm_grad"] = make_dsm_grads()
    elif which == "train":
        manifest["train_steps"] = make_train_steps()
    elif which == "layers":
        manifest["layer_kat"] = make_layer_kat()
    elif which == "segs":
        manifest["segments"] = make_segments()
    elif which == "segs_ref":
        manifest.setdefault("segments", {}).update(make_segments_s4_ref())
    elif which == "mmd":
        manifest["c1_mmd"] = make_mmd_reference()
        print(manifest["c1_mmd"])
    elif which == "nspdk":
        manifest["nspdk"] = make_nspdk_reference()
        print(manifest["nspdk"])
    elif which == "nspdk_symbols":
        return nspdk_symbol_run(sys.argv[2])
    elif which == "traj_new":
        manifest["traj"].update(make_traj(missing_only=True))
    else:
        manifest["c1"] = make_c1_stats()
        print(manifest["c1"])
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    only(sys.argv[1]) if sys.argv[1:2] in (["decode"], ["c1"], ["dsm"], ["mol"], ["dsmgrad"], ["train"], ["traj_new"], ["segs"], ["segs_ref"], ["layers"], ["mmd"], ["nspdk"], ["nspdk_symbols"]) else main()
