"""Per-phase critical-path cycles of fused_layers_kernel (GDSS_PHASE_PROF=1): thread 0 of every CTA accumulates the
clock cycles between phase boundaries into workspace region 'meta' (+128 bytes)."""
import os, sys
os.environ["GDSS_PHASE_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import WORKLOADS, make_flags, load_weights
from gdss_b200.engine import Engine, make_sampler_desc
from gdss_b200.schedule import build_table
from tests.gpu_common import sde_from
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "zinc250k_v2"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2960
sdx, px, sda, pa = load_weights(wl)
flags, counts = make_flags(wl, B, "ragged")
sx, sa = sde_from(wl["sde_x"], 1000), sde_from(wl["sde_adj"], 1000)
table = build_table(sx, sa, wl["predictor"], wl["corrector"], wl["snr"], wl["scale_eps"], 1e-4)
sdesc = make_sampler_desc(sx, sa, wl["predictor"], wl["corrector"], wl["snr"], wl["scale_eps"], 1e-4, 1, False, True)
eng = Engine(sdx, sda, wl["N"], "cuda")
fl = flags.cuda()
x, adj, xm, am = eng.sample(sdesc, table, fl, n_iters=2, seed=1, use_graph=False)
torch.cuda.synchronize()
prof = eng.region(B, "meta", (32,), torch.int64)
prof[16:].zero_()
iters = 3
x, adj, xm, am = eng.sample(sdesc, table, fl, x=x, adj=adj, first_step=2, n_iters=iters, seed=1, use_graph=False)
torch.cuda.synchronize()
p = prof[16:32].cpu().numpy().astype(np.float64)
names = ["stage+pow", "degrees", "A.x", "chain: wait for slowest warp (or QKV, first layer)", "nodeL1+maps", "edge/node MLP: wait", "final X MLP",
         "edge MLP (warp 0)", "channel chain (warp 0)"]
evals = iters * wl["evals"] * B
tot = p[:9].sum()
print(f"graph-evals {evals}, mean cycles per graph-eval {tot / evals:.0f}")
for nme, v in zip(names, p[:9]):
    print(f"  {nme:52s} {v / evals:9.0f} cycles  {v / tot * 100:5.1f}%")
