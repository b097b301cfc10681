"""One process per GPU: rank bookkeeping, work sharding and the one small collective the path has.

Parameter sets (tempering replicas, dense-sampling points) are independent, so ranks take disjoint
contiguous shards and there is no collective on the data path.  The only exchange is the
all-gather of per-chain log-likelihood rows that replica-exchange swaps need
(metropolis.py:204-261 of the reference does this with MPI send/recv pairs).  PyTorch is used for
the plumbing only: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import numpy as np


class Comm:
    """Thin wrapper over torch.distributed (or nothing, when world == 1)."""

    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.backend = None
        self.device = device
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if not dist.is_initialized():
                if backend is None:
                    backend = "nccl" if torch.cuda.is_available() else "gloo"
                if backend == "nccl":
                    torch.cuda.set_device(self.local_rank)
                    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
                else:
                    dist.init_process_group(backend=backend)
            self.dist = dist
            self.backend = dist.get_backend()
            self.torch = torch

    def _dev(self):
        return f"cuda:{self.local_rank}" if self.backend == "nccl" else "cpu"

    def shard(self, n):
        """Contiguous block [lo, hi) of n items owned by this rank (sizes differ by at most one)."""
        return shard_range(n, self.rank, self.world)

    def allgather_rows(self, local, n_total):
        """Concatenate every rank's rows (rank order).  local: [n_local, ...] float64."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.dist is None:
            return local
        torch = self.torch
        tail = local.shape[1:]
        counts = [shard_range(n_total, r, self.world) for r in range(self.world)]
        width = max(hi - lo for lo, hi in counts)
        pad = np.zeros((width,) + tail)
        pad[:local.shape[0]] = local
        send = torch.from_numpy(pad).to(self._dev())
        recv = [torch.empty_like(send) for _ in range(self.world)]
        self.dist.all_gather(recv, send)
        parts = [recv[r][:counts[r][1] - counts[r][0]].cpu().numpy() for r in range(self.world)]
        return np.concatenate(parts, axis=0)

    def allgather_device_rows(self, dev_ptr, n_local, width, n_total):
        """All-gather rows that already sit in this rank's HBM (a raw device pointer from the CUDA
        library): device-to-device over NVLink, then ONE copy of the gathered table to pinned host
        memory.  Returns [n_total, width] float64 on the host."""
        torch = self.torch
        dev = self._dev()
        counts = [shard_range(n_total, r, self.world) for r in range(self.world)]
        most = max(hi - lo for lo, hi in counts)
        key = (most, width)
        if getattr(self, "_gather_key", None) != key:
            self._send = torch.zeros((most, width), dtype=torch.float64, device=dev)
            self._recv = torch.empty((self.world * most, width), dtype=torch.float64, device=dev)
            self._host = torch.empty((self.world * most, width), dtype=torch.float64).pin_memory()
            self._gather_key = key
        view = torch.as_tensor(_DeviceRows(dev_ptr, (n_local, width)), device=dev)
        self._send[:n_local].copy_(view)
        self.dist.all_gather_into_tensor(self._recv, self._send)
        self._host.copy_(self._recv, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = self._host.numpy().reshape(self.world, most, width)
        if all(hi - lo == most for lo, hi in counts):
            return out.reshape(n_total, width).copy()
        return np.concatenate([out[r, :counts[r][1] - counts[r][0]] for r in range(self.world)], axis=0)

    def broadcast_array(self, x, src=0):
        """Rank `src`'s float64 array on every rank (same shape everywhere)."""
        if self.dist is None:
            return x
        t = self.torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self._dev())
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy()

    def max(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        if self.dist is not None:
            if self.backend == "nccl":
                self.dist.barrier(device_ids=[self.local_rank])
            else:
                self.dist.barrier()


class _DeviceRows:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def shard_range(n, rank, world):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
