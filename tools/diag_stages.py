"""Stage-by-stage comparison of the CUDA score networks with the CPU oracle (debugging aid;
reads the named workspace regions).  usage: python tools/diag_stages.py [ckpt ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gdss_oracle as O                      # noqa: E402
from oracle.cases import load_golden_ckpt                # noqa: E402
from gdss_b200.engine import Engine                      # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def offdiag(t):
    n = t.shape[-1]
    return t * (1 - torch.eye(n))


def main(names):
    for name in names:
        ck = load_golden_ckpt(name)
        px, pa = ck["params_x"], ck["params_adj"]
        N, F = pa["max_node_num"], pa["max_feat_num"]
        B = 4 if N < 200 else 2
        x, adj, flags = O.kat_inputs(N, F, B)
        sda, sdx = ck["adj_state_dict"], ck["x_state_dict"]
        with torch.no_grad():
            adjc = O.pow_stack(adj, pa["c_init"])
            stacks = [adjc]
            h = x
            xs_list = []
            for l in range(pa["num_layers"]):
                h, adjc = O.attention_layer(sda, f"layers.{l}", h, adjc, flags, pa["num_linears"], 4)
                stacks.append(adjc)
                xs_list.append(h)
            cat = torch.cat(stacks, dim=1)
            ref_a = O.run_model(sda, pa, x, adj, flags)
            ref_x = O.run_model(sdx, px, x, adj, flags)
        # A network only first (workspace is shared between the two networks)
        eng = Engine(None, sda, N, "cuda")
        _, sa = eng.score(x.cuda(), adj.cuda(), flags.cuda(), inputs_masked=True)
        torch.cuda.synchronize()
        fdim = cat.shape[1]
        st = eng.region(B, "stack_a", (B, fdim, N, N)).cpu()
        print(f"== {name}: N={N} F={F} fdim={fdim}")
        for c in range(fdim):
            print(f"   plane {c:2d}: rel {rel(offdiag(st[:, c]), offdiag(cat[:, c])):.3e}")
        L = pa["num_layers"]
        nh = pa["nhid"]
        for l in range(L - 1):
            buf = eng.region(B, "xa0" if l % 2 == 0 else "xa1", (B, N, nh)).cpu()
            if l >= L - 3:      # only the last writes of each ping-pong buffer survive
                print(f"   x_out layer {l}: rel {rel(buf, xs_list[l]):.3e}")
        print(f"   score_adj: rel {rel(sa, ref_a):.3e}")
        eng2 = Engine(sdx, None, N, "cuda")
        sx, _ = eng2.score(x.cuda(), adj.cuda(), flags.cuda(), inputs_masked=True)
        print(f"   score_x:   rel {rel(sx, ref_x):.3e}")


if __name__ == "__main__":
    main(sys.argv[1:] or ["qm9", "zinc250k_v2", "community_small"])
