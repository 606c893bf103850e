"""tools/err_probe.py : measured rel-L2 errors of the CUDA path against the reference-generated fixtures (raw scores of
every checkpoint; final states + thresholded mismatches of the full 1000-step runs).  GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import gdss_oracle as O
from oracle.cases import CKPTS, FULL_CASES, load_golden_ckpt, load_kat
from tests.gpu_common import rel, injected_noise, case_setup
from gdss_b200.engine import Engine
from gdss_b200.graph_utils import quantize_mol, quantize

for name in CKPTS:
    ck, kat = load_golden_ckpt(name), load_kat(name)
    B, N, F = int(kat["B"]), int(kat["N"]), int(kat["F"])
    x, adj, flags = O.kat_inputs(N, F, B)
    eng = Engine(ck["x_state_dict"], ck["adj_state_dict"], N, "cuda")
    sx, sa = eng.score(x.cuda(), adj.cuda(), flags.cuda())
    print(f"score {name:16s} x {rel(sx, torch.from_numpy(kat['score_x'])):.2e} adj {rel(sa, torch.from_numpy(kat['score_adj'])):.2e}")
for c in FULL_CASES:
    name = c["name"]
    case, z, weights, (N, F, B), flags, table, sdesc, eng = case_setup(name)
    iters = case["scales"]
    px, pa, ix, ia = injected_noise(case, B, N, F, iters)
    it = iter([px, pa])
    x0, a0 = O.prior((B, N, F), (B, N, N), flags, lambda shape: next(it))
    x, adj, xm, am = eng.sample(sdesc, table, flags, x=x0, adj=a0, n_iters=iters, inj_x=ix, inj_adj=ia, use_graph=True)
    gx, ga = torch.from_numpy(z["x"]), torch.from_numpy(z["adj"])
    mism = int((quantize_mol(am) != O.quantize_mol(ga)).sum()) + int((quantize(am).cpu() != O.quantize(ga)).sum()) + int(((xm > 0.5).cpu() != (gx > 0.5)).sum())
    print(f"full  {name:32s} x {rel(xm, gx):.2e} adj {rel(am, ga):.2e} thresholded mismatches {mism}")
