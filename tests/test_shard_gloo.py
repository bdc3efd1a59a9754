"""Host-side logic of the multi-GPU path on CPU: vertex ranges and the weight broadcast over gloo (world_size 2).
The evaluator in these tests is the CPU oracle (tests may use it); on the GPUs the same plumbing moves the
library's device buffers over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facedeform_b200 import shard, synth


def test_vertex_ranges_tile_exactly():
    for V in (0, 1, 7, 100_000, 16_000_000):
        for G in (1, 2, 3, 4, 8):
            ranges = [shard.vertex_range(V, r, G) for r in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == V
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.vertex_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fd_oracle as o
    rig = synth.control_rig(40)
    deform = synth.deformed_rig(rig, 3)
    mesh = synth.face_mesh(V, topology=False)
    p = o.make_params(model=1, term=0, kernel=0, radius=2 * rig.spacing, **{"lambda": 0.0})
    n = 40 + 4
    W = torch.zeros((n, 9), dtype=torch.float64)
    rad = torch.zeros(40, dtype=torch.float64)
    if rank == 0:                                   # only the root factors and solves
        st, r, w = o.fit(p, rig.rest, deform)
        assert st == 1
        W.copy_(torch.from_numpy(w))
        rad.copy_(torch.from_numpy(r))
    shard.broadcast_block(W, 0)                     # the one exchange step of the path
    shard.broadcast_block(rad, 0)
    b, e = shard.vertex_range(V, rank, world)
    out, fall = o.evaluate(p, rig.rest, rad.numpy(), W.numpy(), mesh.P[b:e])
    gathered = [None] * world
    dist.all_gather_object(gathered, (b, e, out, fall))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded_over_gloo():
    from oracle import fd_oracle as o
    V, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    gathered = q.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    rig = synth.control_rig(40)
    deform = synth.deformed_rig(rig, 3)
    mesh = synth.face_mesh(V, topology=False)
    p = o.make_params(model=1, term=0, kernel=0, radius=2 * rig.spacing, **{"lambda": 0.0})
    st, rad, W = o.fit(p, rig.rest, deform)
    ref, rfall = o.evaluate(p, rig.rest, rad, W, mesh.P)
    gathered.sort(key=lambda t: t[0])
    out = np.concatenate([g[2] for g in gathered], axis=1)       # concatenation in rank order = vertex order
    fall = np.concatenate([g[3] for g in gathered])
    assert np.array_equal(out, ref) and np.array_equal(fall, rfall)   # bit-identical to the unsharded run


def test_weights_mode_policy():
    """small systems are solved on every rank (no exchange step), large ones by the root + broadcast."""
    from facedeform_b200 import shard
    assert shard.weights_mode(256) == "replicated" and shard.weights_mode(1024) == "replicated"
    assert shard.weights_mode(2048) == "broadcast" and shard.weights_mode(4096) == "broadcast"
    assert shard.weights_mode(256, "broadcast") == "broadcast" and shard.weights_mode(8192, "replicated") == "replicated"
    import pytest
    with pytest.raises(ValueError):
        shard.weights_mode(10, "gather")
