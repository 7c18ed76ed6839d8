"""Under torchrun on N GPUs of one node: the peer-memory all-reduce (rsrgan_b200/peer.py, csrc/peer_allreduce.cu)
against NCCL on the same data -- bit-identical across ranks, equal to NCCL's sum within fp32 rounding (bit-exact at
N = 2), repeatable inside a CUDA graph -- and the time of both (CUDA events, max over ranks).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/gpu_peer_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops, peer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = ops.Handle(local, "f16")
sizes = [2_449_408, 3_200_000 + 1024 * 37, 4096]            # cfg-2's generator / discriminator stores, a tiny one
comm = peer.try_create(h, dist, sizes)
assert comm is not None, "peer all-reduce not available"
res = {"world": world}
g = torch.Generator(device="cuda").manual_seed(100 + rank)
for i, n in enumerate(sizes):
    buf = comm.buffer(i)
    for it in range(3):
        x = torch.randn(n, device="cuda", generator=g) * (1 + rank)
        buf.copy_(x)
        ref = x.clone()
        dist.all_reduce(ref)
        comm.all_reduce(buf)
        torch.cuda.synchronize()
        err = float((buf - ref).abs().max() / ref.abs().max())
        assert err < 1e-6, (i, it, err)
        if world == 2:
            assert torch.equal(buf, ref)
        lo, hi = buf.clone(), buf.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks differ"
# inside a CUDA graph, replayed
buf = comm.buffer(1)
x = torch.randn(sizes[1], device="cuda", generator=g)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
gr = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    gr.capture_begin(capture_error_mode="thread_local")
    buf.copy_(x)
    comm.all_reduce(buf)
    gr.capture_end()
torch.cuda.current_stream().wait_stream(s)
ref = x.clone()
dist.all_reduce(ref)
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
assert float((buf - ref).abs().max() / ref.abs().max()) < 1e-6


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for i, n in enumerate(sizes):
    buf = comm.buffer(i)
    other = torch.zeros(n, device="cuda")
    res["n%d" % n] = {"peer_us": timeit(lambda: comm.all_reduce(buf)), "nccl_us": timeit(lambda: dist.all_reduce(other)),
                      "bytes": 4 * n}
assert comm.error() == 0
torch.cuda.synchronize()
dist.barrier()
comm.close()
if rank == 0:
    print(json.dumps(res))
    print("PEER_CHECK_OK")
dist.destroy_process_group()
