"""Vertex-range sharding across GPUs.

Every vertex is independent given the weights (the reference loop SOP_FaceDeform.cpp:404-439 has no cross-vertex
dependence), so the only exchange step is the broadcast of the solved weights from the rank that factored the
system; results stay sharded (C5's 192 GB output cannot land on one GPU) or are gathered by the caller.
One process per GPU; torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations


def vertex_range(n_vtx: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous range [begin, end) of rank `rank`: concatenating the ranges in rank order preserves vertex order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_vtx), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


# Up to this many control points every rank factors and solves the (small) system itself instead of waiting for the
# root's broadcast: the fit costs the same on every rank and runs concurrently, so the step loses nothing and the
# exchange step (NCCL launch + 1.5 MB over NVLink + the wait for the root) disappears -- SURVEY.md section 8e's
# "alternative without the broadcast (redundant solve per rank)", valid while 2/3 N^3 << V N 3F / G.
REPLICATE_MAX_CTRL = 1024


def weights_mode(n_ctrl: int, mode: str = "auto") -> str:
    """how the ranks of a vertex-sharded evaluation get their weights: "broadcast" (root fits and solves, NCCL
    broadcast of the weight block) or "replicated" (every rank fits and solves; no collective).  Both give bit-identical
    weights (same deterministic kernels on the same inputs)."""
    if mode not in ("auto", "broadcast", "replicated"):
        raise ValueError(mode)
    if mode != "auto":
        return mode
    return "replicated" if int(n_ctrl) <= REPLICATE_MAX_CTRL else "broadcast"


class _CudaBlock:
    """zero-copy view of a device allocation owned by libfacedeform_gpu.so (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None,
        }


def device_view(ptr: int, nbytes: int):
    import torch
    return torch.as_tensor(_CudaBlock(ptr, nbytes), device="cuda")


def broadcast_block(t, src: int = 0, group=None):
    """broadcast a tensor in place from `src` (NCCL on CUDA tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


def broadcast_block_tree(t, src: int = 0, group=None):
    """binomial-tree broadcast with point-to-point sends: log2(world) hops instead of the ring's world - 1."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return t
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return t
    rel = (rank - src) % world
    step = 1
    while step < world:
        step *= 2
    step //= 2
    # top-down: in the round with distance `step`, ranks that already hold the data (rel % (2 step) == 0) send to rel + step
    while step >= 1:
        if rel % (2 * step) == 0 and rel + step < world:
            dist.send(t, dst=(rel + step + src) % world, group=group)
        elif rel % (2 * step) == step:
            dist.recv(t, src=(rel - step + src) % world, group=group)
        step //= 2
    return t


def broadcast_model(model, src: int = 0, group=None, shared_stream: bool = False):
    """Root: `model` is fitted + solved.  Others: `model` is a receiver (Context.receiver).  After this call every
    rank can evaluate: the FP64 weight block and the radii travel, the evaluation tables are rebuilt locally.

    shared_stream=True: the library's ctx stream IS torch's current stream, so the collective is ordered after the
    solve and before the evaluation by stream order alone and no host synchronisation is needed."""
    import torch
    import torch.distributed as dist
    wp, wb = model.weights_dev()
    rp, rb = model.radii_dev()
    if not shared_stream:
        model.ctx.synchronize()  # the library's stream and the collective's stream are different streams
    broadcast_block(device_view(wp, wb), src, group)
    broadcast_block(device_view(rp, rb), src, group)
    if not shared_stream:
        torch.cuda.current_stream().synchronize()
    if dist.is_initialized() and dist.get_rank(group) != src:
        model.commit_weights()
    return model
