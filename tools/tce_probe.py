"""Per-graph difference between the tcgen05 and the mma.sync per-edge MLP of the per-graph kernel (raw scores of one
evaluation), for graphs of chosen node counts.  GDSS_TCEDGE is read at context creation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import WORKLOADS, load_weights
from gdss_b200.engine import Engine
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "zinc250k_v2"]
N, F = wl["N"], wl["F"]
counts = [c for c in [5, 8, 12, 16, 17, 20, 23, 24, 30, 32, 33, 36, 38] if c <= N]
B = len(counts)
g = torch.Generator().manual_seed(0)
flags = torch.zeros(B, N)
for b, c in enumerate(counts):
    flags[b, :c] = 1
x = torch.randn(B, N, F, generator=g) * flags[:, :, None]
a = torch.randn(B, N, N, generator=g).triu(1)
a = (a + a.transpose(1, 2)) * flags[:, :, None] * flags[:, None, :]
sdx, px, sda, pa = load_weights(wl)
outs = {}
for mode in ("0", "1"):
    os.environ["GDSS_TCEDGE"] = "1" if mode == "0" else "0"     # outs["0"] = tcgen05 path, outs["1"] = mma.sync path
    eng = Engine(sdx, sda, N, "cuda")
    sx, sa = eng.score(x, a, flags, inputs_masked=True)
    torch.cuda.synchronize()
    outs[mode] = (sx.double().cpu(), sa.double().cpu())
for b, c in enumerate(counts):
    rx = float((outs["0"][0][b] - outs["1"][0][b]).norm() / outs["1"][0][b].norm())
    ra = float((outs["0"][1][b] - outs["1"][1][b]).norm() / outs["1"][1][b].norm())
    d = (outs["0"][1][b] - outs["1"][1][b]).abs()
    ij = divmod(int(d.argmax()), N)
    print(f"n={c:3d} P={c*(c-1)//2:4d}  rel x {rx:.2e}  rel adj {ra:.2e}  max|d adj| {float(d.max()):.3e} at {ij}")
