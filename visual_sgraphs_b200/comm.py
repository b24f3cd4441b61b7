"""vsg_comm over the C ABI: the NCCL communicator of the two sharded matcher paths (include/vsg_cuda.h).
The 128-byte id is created on rank 0 and handed to the other ranks by whatever the host application has —
here torch.distributed's broadcast (plumbing only)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

ID_BYTES = 128


class Comm:
    def __init__(self, unique_id, nranks, rank, device):
        self._L = _lib.load()
        self._h = C.c_void_p()
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == ID_BYTES
        check(self._L.vsg_comm_create(ptr(uid), int(nranks), int(rank), int(device), C.byref(self._h)))
        self.rank, self.size = rank, nranks

    @staticmethod
    def unique_id():
        uid = np.zeros(ID_BYTES, np.uint8)
        check(_lib.load().vsg_comm_unique_id(ptr(uid)))
        return uid

    @staticmethod
    def nccl_version():
        return _lib.load().vsg_comm_nccl_version()

    @classmethod
    def from_torch_distributed(cls, dist, device):
        """One communicator spanning the ranks of an initialised torch.distributed process group."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = torch.from_numpy(cls.unique_id() if rank == 0 else np.zeros(ID_BYTES, np.uint8))
        if world > 1:
            on_gpu = dist.get_backend() == "nccl"
            t = uid.cuda(device) if on_gpu else uid
            dist.broadcast(t, 0)
            uid = t.cpu()
        return cls(uid.numpy(), world, rank, device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsg_comm_destroy(self._h)
            self._h = None

    __del__ = close
