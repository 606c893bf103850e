import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from oracle.cases import load_traj, NumpyDraws, case_weights
from oracle import gdss_oracle as O
torch.set_num_threads(8)
case, z = load_traj("zinc_v2_rev_lang_full_b32")
flags = torch.from_numpy(z["flags"])
sdx, px, sda, pa = case_weights(case)
N, F, B = pa["max_node_num"], pa["max_feat_num"], case["B"]
sx = O.SDESpec(*case["sde_x"], case["scales"]); sa = O.SDESpec(*case["sde_adj"], case["scales"])
base = NumpyDraws(case["seed"])
cnt = [0]
def draw(shape):
    t = base(shape)
    cnt[0] += 1
    if cnt[0] == 2:
        t = t * (1.0 + 2.0e-7)      # ~1-2 ulp perturbation of the adjacency prior
    return t
t0=time.time()
with torch.no_grad():
    x, adj, n = O.pc_sample(sx, sa, sdx, px, sda, pa, (B,N,F), (B,N,N), flags, draw, predictor=case["predictor"], corrector=case["corrector"], snr=case["snr"], scale_eps=case["scale_eps"], denoise=True, eps=1e-4)
gx, ga = torch.from_numpy(z["x"]), torch.from_numpy(z["adj"])
rel=lambda a,b: float((a-b).norm()/b.norm())
print("perturbed-prior oracle vs reference: rel x %.3e adj %.3e (%.0fs)"%(rel(x,gx), rel(adj,ga), time.time()-t0))
print("per graph adj rel:", ["%.1e"%float((adj[b]-ga[b]).norm()/ga[b].norm()) for b in range(B)])
print("quantize_mol mismatches:", int((O.quantize_mol(adj)!=O.quantize_mol(ga)).sum()))
