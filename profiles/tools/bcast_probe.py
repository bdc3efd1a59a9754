"""Development probe (torchrun): latency of moving the C2 weight block (1.5 MB) from rank 0 to all ranks."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from facedeform_b200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.zeros(1_500_000, dtype=torch.uint8, device="cuda")
small = torch.zeros(2048, dtype=torch.uint8, device="cuda")


def timeit(fn, n=50):
    for _ in range(5):
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
    ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


a = timeit(lambda: dist.broadcast(t, src=0))
b = timeit(lambda: dist.broadcast(small, src=0))
c = timeit(lambda: shard.broadcast_block_tree(t, 0))
if rank == 0:
    print(f"world={world}: nccl broadcast 1.5MB {a * 1e3:.1f} us, 2KB {b * 1e3:.1f} us, tree send/recv 1.5MB {c * 1e3:.1f} us")
dist.destroy_process_group()
