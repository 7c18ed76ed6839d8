"""csrc/peer_allreduce.cu: average_gradients (utils/ops.py:343-376) as one kernel over peer memory.

On ONE GPU the protocol is exercised with emulated ranks: `world` blocks from rsr_peer_alloc in this process, one
kernel per "rank" on its own stream (all co-resident: max_blocks x world <= 148 CTAs), exchanging through the same
flags and slices the real ranks use over NVLink.  Integer-exact bar: the sum is taken in rank order 0..world-1 in fp32,
so every buffer must equal ((b0 + b1) + b2) + ... BIT FOR BIT, on every rank.  With two or more GPUs the real thing
(CUDA IPC between processes, under torchrun) is run by scripts/gpu_peer_check.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HEADER = 16384


@pytest.fixture(scope="module")
def h():
    from rsrgan_b200 import ops
    hd = ops.Handle(0, "f16")
    yield hd
    hd.close()


def _span(ptr, n, dev):
    from rsrgan_b200.peer import _DeviceSpan
    return torch.as_tensor(_DeviceSpan(ptr, n, None), device=dev)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n", [4, 1000, 12 * 1024 + 4, 3_200_000])
def test_emulated_ranks_sum_bit_exact_and_repeatable(h, world, n):
    dev, rng = h.device, np.random.default_rng(world * 7 + n)
    blocks, bufs = [], []
    for r in range(world):
        blk, _ = h.peer_alloc(4 * n)
        blocks.append(blk)
        bufs.append(_span(blk + HEADER, n, dev))
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    try:
        for it in range(3):                                   # monotonic barrier counters: nothing is reset between calls
            src = [torch.tensor((rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3)).astype(np.float32), device=dev)
                   for _ in range(world)]
            want = src[0].clone()
            for r in range(1, world):
                want += src[r]                                # fp32, rank order
            for r in range(world):
                bufs[r].copy_(src[r])
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    h.peer_allreduce(blocks, r, HEADER, n, max_blocks=16)
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(bufs[r], want), (it, r)
                assert h.peer_error(blocks[r]) == 0
    finally:
        torch.cuda.synchronize()
        for blk in blocks:
            h.peer_free(blk)


def test_argument_errors(h):
    from rsrgan_b200._lib import RsrError
    blk, ipc = h.peer_alloc(64)
    try:
        assert len(ipc) == 64
        with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # world size the kernels are not built for
            h.peer_allreduce([blk] * 3, 0, HEADER, 16)
        with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # buffer inside the header
            h.peer_allreduce([blk] * 2, 0, 0, 16)
        with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # not whole float4s
            h.peer_allreduce([blk] * 2, 0, HEADER, 6)
        with pytest.raises(RsrError, match="RSR_E_ARG"):
            h.peer_allreduce([blk, blk], 2, HEADER, 16)
        h.peer_allreduce([blk], 0, HEADER, 16)                        # world 1: nothing to do
    finally:
        h.peer_free(blk)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_two_processes_over_cuda_ipc():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "scripts", "gpu_peer_check.py")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "PEER_CHECK_OK" in out.stdout
