"""The data-parallel path's one collective through the C ABI (csrc/comm.cu): a sum all-reduce of the flat fp64 gradient
buffer on an npgp communicator (NCCL bound at run time inside libnpgp.so).  torch.distributed is used once, to ship the
128-byte rendezvous token from rank 0 to the other ranks."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr, stream


class NpgpComm:
    """comm = NpgpComm(rank, world, device); comm(t) all-reduces the fp64 CUDA tensor t in place on the current stream."""

    def __init__(self, rank: int, world: int, device: torch.device):
        import torch.distributed as dist
        token = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            check(lib().npgp_comm_unique_id(C.cast(buf, C.c_void_p)), "npgp_comm_unique_id")
            token = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if world > 1:
            on_gpu = dist.get_backend() == "nccl"
            t = token.to(device) if on_gpu else token
            dist.broadcast(t, src=0)
            token = t.cpu()
        raw = C.create_string_buffer(bytes(token.tolist()), 128)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().npgp_comm_create(C.byref(handle), raw, world, rank), "npgp_comm_create")
        self.handle, self.rank, self.world, self.device = handle, rank, world, device

    def all_reduce(self, t: torch.Tensor):
        assert t.dtype == torch.float64 and t.is_contiguous()
        check(lib().npgp_allreduce_f64(self.handle, ptr(t), t.numel(), stream()), "npgp_allreduce_f64")
        return t

    __call__ = all_reduce

    def destroy(self):
        if self.handle:
            check(lib().npgp_comm_destroy(self.handle), "npgp_comm_destroy")
            self.handle = None
