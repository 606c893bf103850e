import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import stats as S
from oracle.cases import load_golden_ckpt, GOLD
from gdss_b200.solver import get_pc_sampler
from gdss_b200.sde import VPSDE
from gdss_b200.graph_utils import quantize
z = np.load(os.path.join(GOLD, "c1_community_reference.npz"))
counts, ref, train, tc = z["counts"].astype(int), z["samples"], z["train"], z["train_counts"].astype(int)
ck = load_golden_ckpt("community_small")
cnt = np.tile(counts, 10); B, N, F = len(cnt), 20, 10
flags = (torch.arange(N)[None, :] < torch.from_numpy(cnt)[:, None]).float()
fn = get_pc_sampler(VPSDE(.1, 1., 1000), VPSDE(.1, 1., 1000), (B, N, F), (B, N, N), predictor="Euler", corrector="Langevin", snr=.05, scale_eps=.7, n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-4, device="cuda")
torch.manual_seed(11)
import time; t0=time.time()
x, adj, n = fn(ck["x_state_dict"], ck["adj_state_dict"], flags); torch.cuda.synchronize(); t=time.time()-t0
q = quantize(adj).cpu().numpy().astype(np.int8)
e_ref, e = S.edge_counts(ref), S.edge_counts(q)
valid = lambda c: (c * (c - 1) / 2).sum()
ht = S.degree_histograms(train, tc)
print("time", t, "edges ours %.2f +- %.2f ref %.2f +- %.2f" % (e.mean(), e.std(), e_ref.mean(), e_ref.std()), "density", e.sum()/valid(cnt), e_ref.sum()/valid(counts),
      "mmd ours", S.degree_mmd(ht, S.degree_histograms(q[:300], cnt[:300])), "ref", S.degree_mmd(ht, S.degree_histograms(ref, counts)))
