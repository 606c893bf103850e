"""One process per GPU: contiguous batch shards, the library's own NCCL communicator for the
per-step 4-float Langevin all-reduce and the final gather."""
import ctypes as C
import os

import torch
import torch.distributed as dist

from ._lib import check, lib


def shard_bounds(total, world, rank):
    """Contiguous shard [lo, hi) of `total` graphs for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardInfo:
    """world/rank + (on CUDA) an NCCL communicator created from an id broadcast through
    torch.distributed (any backend: gloo on CPU tests, nccl on the GPU box)."""

    def __init__(self, world=None, rank=None, create_comm=True, device=None):
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        self.comm = None
        self.device = device
        if create_comm and self.world > 1:
            self.comm = self._create_comm()

    def bounds(self, total):
        return shard_bounds(total, self.world, self.rank)

    def exchange_id(self, make_id):
        """rank 0 calls make_id() -> 128 bytes; everyone gets the same bytes."""
        buf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = torch.frombuffer(bytearray(make_id()), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            b = buf.to(dev)
            dist.broadcast(b, 0)
            buf = b.cpu()
        else:
            dist.broadcast(buf, 0)
        return bytes(buf.tolist())

    def _create_comm(self):
        L = lib()

        def make():
            raw = C.create_string_buffer(128)
            check(L.gdss_nccl_unique_id(raw), "gdss_nccl_unique_id")
            return raw.raw

        idb = self.exchange_id(make)
        comm = C.c_void_p()
        check(L.gdss_nccl_comm_create(C.create_string_buffer(idb, 128), self.world, self.rank, C.byref(comm)),
              "gdss_nccl_comm_create")
        return comm

    def gather(self, t, total):
        """All-gather the contiguous shards along dim 0 -> full [total, ...] tensor on every rank."""
        if self.world == 1:
            return t
        sizes = [shard_bounds(total, self.world, r) for r in range(self.world)]
        mx = max(hi - lo for lo, hi in sizes)
        per = t[0].numel()
        send = torch.zeros(mx * per, device=t.device)
        send[: t.numel()] = t.reshape(-1)
        recv = torch.empty(self.world * mx * per, device=t.device)
        check(lib().gdss_nccl_allgather(self.comm, send.data_ptr(), recv.data_ptr(), mx * per,
                                        torch.cuda.current_stream(t.device).cuda_stream), "gdss_nccl_allgather")
        parts = [recv[r * mx * per: r * mx * per + (hi - lo) * per] for r, (lo, hi) in enumerate(sizes)]
        return torch.cat(parts).view(total, *t.shape[1:])

    def close(self):
        if self.comm is not None:
            lib().gdss_nccl_comm_destroy(self.comm)
            self.comm = None


def init_from_env():
    """torchrun environment -> (ShardInfo, device).  MASTER_ADDR should be 127.0.0.1 on one node."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    return ShardInfo(world, rank, create_comm=torch.cuda.is_available(), device=dev), dev
