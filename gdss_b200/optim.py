"""Optimizer half of the reference's training step (trainer.py:57-79; utils/loader.py:51-62 builds
``torch.optim.Adam(model.parameters(), lr, weight_decay)``, utils/ema.py the ExponentialMovingAverage) on ONE flat fp32
parameter vector per network:

    clip_grad_norm_(params, grad_norm) -> Adam.step() -> ema.update()

as two CUDA launches (gdss_adam_ema_step).  Data parallel: ``allreduce_grads`` sums the flat gradient over the ranks with
NCCL first (the clip must see the global gradient); ``grad_scale = 1 / world`` turns the sum into the batch mean.

The gradients come from ``gdss_dsm_loss_backward`` (``gdss_b200.trainer.DSMTrainStep`` wires the three calls together) or from
the reference's own autograd; ``advance`` / ``enqueue_dev`` are the same step for a CUDA-graph replay (per-step scalars on the device).
"""
import ctypes as C

import torch

from ._lib import check, lib, require_cuda


def flatten_state(sd):
    """state_dict (parameters() order == state_dict order for the score networks: they have no buffers) -> flat fp32
    vector + the (name, shape) list to unflatten it."""
    metas = [(k, tuple(v.shape)) for k, v in sd.items()]
    flat = torch.cat([v.detach().reshape(-1).to(torch.float32) for v in sd.values()])
    return flat, metas


def unflatten_state(flat, metas):
    out, o = {}, 0
    for k, shape in metas:
        n = 1
        for d in shape:
            n *= d
        out[k] = flat[o:o + n].view(shape)
        o += n
    return out


class FusedAdamEMA:
    """Adam moments + EMA shadow of one network's flat parameter vector on a CUDA device.

    lr / weight_decay / betas / eps: torch.optim.Adam's; grad_norm: the trainer's clip (config train.grad_norm, <= 0: none);
    ema_decay: config train.ema (None: no EMA); active: optional uint8 vector, 0 for parameters the loss never reaches
    (grad None in torch, skipped by torch.optim.Adam: no weight decay, no moment update)."""

    def __init__(self, flat_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_norm=0.0, ema_decay=None,
                 active=None):
        require_cuda()
        if flat_params.device.type != "cuda" or flat_params.dtype != torch.float32 or not flat_params.is_contiguous():
            raise ValueError("flat_params must be a contiguous fp32 CUDA tensor")
        self.L = lib()
        self.p = flat_params
        self.m = torch.zeros_like(flat_params)
        self.v = torch.zeros_like(flat_params)
        self.shadow = flat_params.clone() if ema_decay is not None else None        # utils/ema.py:12-13
        self.lr, self.betas, self.eps, self.wd = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.grad_norm, self.ema_decay = float(grad_norm or 0.0), ema_decay
        self.step_count, self.num_updates = 0, 0
        self.active = None
        if active is not None:
            self.active = active.to(flat_params.device, torch.uint8).contiguous()
            if self.active.numel() != flat_params.numel():
                raise ValueError("active mask does not match the parameters")
        self.scratch = torch.zeros(int(self.L.gdss_optim_scratch_floats()), device=flat_params.device)
        self.norm = torch.zeros(1, device=flat_params.device)

    def step(self, flat_grads, grad_scale=1.0):
        """One trainer step; returns the (device) total gradient norm before clipping when clipping is on."""
        if flat_grads.shape != self.p.shape or flat_grads.device != self.p.device:
            raise ValueError("gradient vector does not match the parameters")
        self.step_count += 1
        st = torch.cuda.current_stream(self.p.device).cuda_stream
        with torch.cuda.device(self.p.device):
            check(self.L.gdss_adam_ema_step(self.p.data_ptr(), flat_grads.contiguous().data_ptr(), self.m.data_ptr(),
                                            self.v.data_ptr(), self.shadow.data_ptr() if self.shadow is not None else None,
                                            self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                            self.grad_norm, float(grad_scale), self.step_count,
                                            float(self.ema_decay if self.ema_decay is not None else 0.0), self.num_updates,
                                            self.active.data_ptr() if self.active is not None else None,
                                            self.scratch.data_ptr(), self.norm.data_ptr(), st), "gdss_adam_ema_step")
        if self.shadow is not None:
            self.num_updates += 1
        return self.norm if self.grad_norm > 0 else None

    # ---- the same step for a captured CUDA graph: the launch (`enqueue_dev`, captured once) reads the four per-step scalars from
    # `self.consts`; `advance()` (every step, before the replay) refreshes them from the host-side step count
    def advance(self):
        """Bump the step count and copy {lr, bias corrections, EMA factor} of this step to the device (stream-ordered).
        Every step stages its scalars in its OWN pinned block: the host runs ahead of the device when steps are replayed
        back to back, and a single reused staging buffer would be overwritten with a later step's scalars before the queued
        copy of this step has read it.  (torch's pinned-memory cache recycles a block only after the copies that read it.)"""
        self.step_count += 1
        if not hasattr(self, "consts"):
            self.consts = torch.zeros(4, device=self.p.device)
        host = torch.zeros(4, dtype=torch.float32).pin_memory()
        check(self.L.gdss_adam_step_consts(self.lr, self.betas[0], self.betas[1], self.step_count,
                                           float(self.ema_decay if self.ema_decay is not None else 0.0), self.num_updates,
                                           host.data_ptr()), "gdss_adam_step_consts")
        self.consts.copy_(host, non_blocking=True)
        if self.shadow is not None:
            self.num_updates += 1

    def enqueue_dev(self, flat_grads, grad_scale=1.0):
        if not hasattr(self, "consts"):
            self.consts = torch.zeros(4, device=self.p.device)
        st = torch.cuda.current_stream(self.p.device).cuda_stream
        with torch.cuda.device(self.p.device):
            check(self.L.gdss_adam_ema_step_dev(self.p.data_ptr(), flat_grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                                self.shadow.data_ptr() if self.shadow is not None else None, self.p.numel(),
                                                self.betas[0], self.betas[1], self.eps, self.wd, self.grad_norm, float(grad_scale),
                                                self.consts.data_ptr(), self.active.data_ptr() if self.active is not None else None,
                                                self.scratch.data_ptr(), self.norm.data_ptr(), st), "gdss_adam_ema_step_dev")
        return self.norm if self.grad_norm > 0 else None


def allreduce_grads(shard, flat_grads):
    """In-place NCCL sum of the flat gradient over the ranks of a ShardInfo (gdss_b200.dist); no-op for one rank."""
    if shard is None or shard.world == 1:
        return flat_grads
    st = torch.cuda.current_stream(flat_grads.device).cuda_stream
    check(lib().gdss_nccl_allreduce(shard.comm, flat_grads.data_ptr(), flat_grads.numel(), st), "gdss_nccl_allreduce")
    return flat_grads
