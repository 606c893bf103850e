"""Diagnosis of a full-length injected-noise run: per-graph error of the CUDA sampler against the reference fixture, with and
without dead-chain elimination, and the spread between two CUDA variants (sensitivity of the 1000-step trajectory)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import gdss_oracle as O
from tests.gpu_common import rel, injected_noise, case_setup

name = sys.argv[1] if len(sys.argv) > 1 else "zinc_v2_rev_lang_full_b32"
outs = {}
for tag, env in (("pruned", {}), ("noprune", {"GDSS_NO_PRUNE": "1"}), ("noprune_nomma", {"GDSS_NO_PRUNE": "1", "GDSS_NO_MMA": "1"})):
    os.environ.update(env)
    case, z, weights, (N, F, B), flags, table, sdesc, eng = case_setup(name)
    for k in env:
        del os.environ[k]
    iters = case["scales"]
    px, pa, ix, ia = injected_noise(case, B, N, F, iters)
    it = iter([px, pa])
    x0, a0 = O.prior((B, N, F), (B, N, N), flags, lambda shape: next(it))
    x, adj, xm, am = eng.sample(sdesc, table, flags, x=x0, adj=a0, n_iters=iters, inj_x=ix, inj_adj=ia, use_graph=True)
    outs[tag] = (xm.cpu(), am.cpu())
    ga = torch.from_numpy(z["adj"])
    per = [float((am.cpu()[b] - ga[b]).norm() / ga[b].norm()) for b in range(B)]
    print(tag, "vs reference: x %.2e adj %.2e" % (rel(xm, torch.from_numpy(z["x"])), rel(am, ga)))
    print("   per graph adj:", " ".join("%.0e" % v for v in per))
    print("   quantize_mol mismatches:", int((O.quantize_mol(am.cpu()) != O.quantize_mol(ga)).sum()), "of", ga.numel())
for a, b in (("pruned", "noprune"), ("noprune", "noprune_nomma")):
    print(a, "vs", b, ": x %.2e adj %.2e" % (rel(outs[a][0], outs[b][0]), rel(outs[a][1], outs[b][1])))
