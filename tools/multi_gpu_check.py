"""torchrun --nproc-per-node 2 tools/multi_gpu_check.py : the sharded sampler (4-float NCCL all-reduce of the Langevin
norms every step + final all-gather) returns on every rank what one GPU returns for the whole batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from gdss_b200.dist import init_from_env
from gdss_b200.solver import get_pc_sampler, S4_solver
from gdss_b200.sde import VPSDE, VESDE
from gdss_b200.ckpt import load_ckpt as load_golden_ckpt

shard, dev = init_from_env()
ck = load_golden_ckpt("zinc250k_v2")
B, N, F, steps = 37, 38, 9, 6
g = torch.Generator().manual_seed(5)
flags = torch.zeros(B, N)
for b in range(B):
    flags[b, : int(torch.randint(4, N + 1, (1,), generator=g))] = 1
sx, sa = VPSDE(.1, 1., 1000), VESDE(.2, 1., 1000)
kw = dict(predictor="Reverse", corrector="Langevin", snr=.2, scale_eps=.8, n_steps=1, eps=1e-4, device=str(dev), max_iters=steps)
torch.manual_seed(7)
full = get_pc_sampler(sx, sa, (B, N, F), (B, N, N), **kw)(ck["x_state_dict"], ck["adj_state_dict"], flags)
torch.manual_seed(7)
part = get_pc_sampler(sx, sa, (B, N, F), (B, N, N), shard=shard, **kw)(ck["x_state_dict"], ck["adj_state_dict"], flags)
torch.cuda.synchronize()
ex = float((full[0] - part[0]).abs().max()); ea = float((full[1] - part[1]).abs().max())
rx = float((full[0] - part[0]).norm() / full[0].norm()); ra = float((full[1] - part[1]).norm() / full[1].norm())
ok = rx <= 1e-5 and ra <= 1e-5 and bool(torch.isfinite(part[1]).all())
print(f"rank {shard.rank}/{shard.world}: sharded vs single-GPU  max|dx| {ex:.3e} max|dadj| {ea:.3e}  rel {rx:.2e} {ra:.2e}  {'OK' if ok else 'FAIL'}", flush=True)

# data-parallel gradient exchange of the trainer (gdss_nccl_allreduce through gdss_b200.optim.allreduce_grads) + one fused
# clip / Adam / EMA step with grad_scale = 1 / world: every rank must end with the same parameters as a single process that
# saw the mean gradient
from gdss_b200.optim import FusedAdamEMA, allreduce_grads
n = 85594                                                     # ZINC250k v2: 22 721 + 62 873 parameters
gen = torch.Generator().manual_seed(1)
p0 = torch.randn(n, generator=gen)
gr = [torch.randn(n, generator=gen) for _ in range(shard.world)]
mine = gr[shard.rank].to(dev)
allreduce_grads(shard, mine)
torch.cuda.synchronize()
want = torch.stack(gr).sum(0)
e_sum = float((mine.cpu() - want).abs().max())
opt = FusedAdamEMA(p0.to(dev).contiguous(), lr=1e-3, grad_norm=1.0, ema_decay=0.999)
opt.step(mine, grad_scale=1.0 / shard.world)
ref = torch.nn.Parameter(p0.clone())
ref.grad = want / shard.world
torch.nn.utils.clip_grad_norm_([ref], 1.0)
torch.optim.Adam([ref], lr=1e-3).step()
e_p = float((opt.p.cpu() - ref.detach()).abs().max())
ok2 = e_sum <= 1e-5 and e_p <= 2e-6
print(f"rank {shard.rank}/{shard.world}: gradient all-reduce max err {e_sum:.2e}, parameters after the fused step vs torch {e_p:.2e}  {'OK' if ok2 else 'FAIL'}", flush=True)
t = torch.tensor([0 if (ok and ok2) else 1], device=dev)
dist.all_reduce(t)
shard.close()
dist.destroy_process_group()
sys.exit(int(t.item()) != 0)
