"""Gradient averaging over NVLink peer memory -- the host side of csrc/peer_allreduce.cu.

utils/ops.py:343-376 `average_gradients` is the one exchange step of the data-parallel path
(models/gan_rnn_placeholder.py:144-160: every tower's gradients, averaged, feed both optimizers).  With one process per
GPU on one node the flat gradient buffers of the two networks are placed in ONE cudaMalloc'ed block per rank; the
ranks exchange the blocks' CUDA-IPC handles once (through torch.distributed, at construction) and from then on an
all-reduce is a single kernel launch on the compute stream -- no NCCL call, no graph-segment cut: the whole batch
schedule stays one CUDA graph (GAN_RNN._update).

torch.distributed (NCCL) remains the plumbing: rendezvous, the handle exchange, barriers, and the fallback when the
ranks cannot open each other's memory (different nodes, no peer access, RSR_PEER_ALLREDUCE=0)."""
from __future__ import annotations

import os

import torch


class _DeviceSpan(object):
    """`__cuda_array_interface__` view of a span of device memory torch did not allocate."""

    def __init__(self, ptr, n_floats, owner):
        self.__cuda_array_interface__ = dict(shape=(int(n_floats),), typestr="<f4", data=(int(ptr), False), version=2)
        self._owner = owner


class PeerComm(object):
    """One block per rank: [header | buffer 0 | buffer 1 | ...] (each buffer a flat fp32 gradient store).
    Construct through try_create()."""

    def __init__(self, handle, dist, sizes):
        """sizes: floats per buffer (multiples of 4).  Collective: every rank constructs it at the same point."""
        self.h, self.dist = handle, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world not in (2, 4, 8):
            raise ValueError("peer all-reduce: world size must be 2, 4 or 8")
        self.offsets, off = [], handle.PEER_HEADER_BYTES
        for n in sizes:
            assert n % 4 == 0
            self.offsets.append(off)
            off += ((4 * int(n) + 255) // 256) * 256
        self.sizes = [int(n) for n in sizes]
        # the collectives below are reached by every rank whatever fails locally; try_create() follows up with a
        # MIN all-reduce of the outcome, which is also the barrier "every rank has mapped every block"
        try:
            self.block, ipc = handle.peer_alloc(off - handle.PEER_HEADER_BYTES)
        except Exception:               # noqa: BLE001
            self.block, ipc = None, None
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (os.uname().nodename, None if ipc is None else bytes(ipc)))
        self.blocks = None
        if any(g[1] is None for g in gathered) or len({g[0] for g in gathered}) != 1:
            if self.block is not None:
                handle.peer_free(self.block)
            raise RuntimeError("peer all-reduce: allocation failed on a rank, or ranks on different nodes")
        opened = []
        try:
            for r in range(self.world):
                opened.append(self.block if r == self.rank else handle.peer_open(gathered[r][1]))
        except Exception:
            for r, b in enumerate(opened):
                if r != self.rank:
                    handle.peer_close(b)
            handle.peer_free(self.block)
            raise
        self.blocks = opened
        self.tensors = [torch.as_tensor(_DeviceSpan(self.block + o, n, self), device=handle.device)
                        for o, n in zip(self.offsets, self.sizes)]
        torch.cuda.synchronize()
        self.calls = 0

    def buffer(self, i):
        return self.tensors[i]

    def all_reduce(self, t):
        """In-place SUM over the ranks of buffer `t` (one of self.tensors), on torch's current stream."""
        for o, n, own in zip(self.offsets, self.sizes, self.tensors):
            if own.data_ptr() == t.data_ptr() and t.numel() == n:
                self.h.peer_allreduce(self.blocks, self.rank, o, n)
                self.calls += 1
                return
        raise ValueError("peer all-reduce: not one of this communicator's buffers")

    def error(self):
        """1 if a barrier of an all-reduce timed out on this rank (a peer died or issued a different call sequence)."""
        return self.h.peer_error(self.block)

    def close(self):
        if self.blocks is None:
            return
        torch.cuda.synchronize()
        self.tensors = None
        for r, b in enumerate(self.blocks):
            if r != self.rank:
                self.h.peer_close(b)
        self.h.peer_free(self.block)
        self.blocks = None


def try_create(handle, dist, sizes):
    """PeerComm, or None (-> NCCL) when it does not apply.  The decision is made collectively so that all ranks
    take the same path."""
    if handle.device.type != "cuda" or os.environ.get("RSR_PEER_ALLREDUCE", "1") == "0":
        return None
    world, rank = dist.get_world_size(), dist.get_rank()
    ok = world in (2, 4, 8)
    if ok:
        dev = handle.device.index
        ok = all(torch.cuda.can_device_access_peer(dev, d) for d in range(torch.cuda.device_count()) if d != dev) \
            and torch.cuda.device_count() >= world
    flag = torch.tensor([1 if ok else 0], device=handle.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return None
    try:
        comm, made = PeerComm(handle, dist, sizes), 1
    except Exception as e:              # noqa: BLE001 -- any rank failing sends every rank to NCCL
        import warnings
        warnings.warn("peer all-reduce unavailable (%s); using NCCL" % (e,))
        comm, made = None, 0
    flag = torch.tensor([made], device=handle.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if comm is not None:
            comm.close()
        return None
    return comm
