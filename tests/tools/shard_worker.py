"""torchrun worker of tests/test_gpu_round2.py::test_torchrun_ranks_match_one_gpu_bit_for_bit: one process per GPU,
rank 0 fits and solves, the weights are broadcast (facedeform_b200.shard.broadcast_model: NCCL over NVLink), every
rank evaluates its vertex range; rank 0 also evaluates everything alone and stores both results."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from facedeform_b200 import Context, make_params, shard, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group(a.backend, device_id=torch.device("cuda", local))
    N, F, V = 300, 48, 20_003
    rig = synth.control_rig(N)
    deform = synth.deformed_rig(rig, F)
    mesh = synth.face_mesh(V, topology=False)
    p = make_params(model=1, radius=2 * rig.spacing, **{"lambda": 0.0})
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(local, stream=stream.cuda_stream)
    d_rest = torch.from_numpy(rig.rest).cuda()
    if rank == 0:
        m = ctx.fit(p, d_rest).solve(torch.from_numpy(deform).cuda())
    else:
        m = ctx.receiver(p, d_rest, F)
    shard.broadcast_model(m, 0, shared_stream=True)
    b, e = shard.vertex_range(V, rank, world)
    out, _ = m.eval(torch.from_numpy(np.ascontiguousarray(mesh.P[b:e])).cuda())
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (b, out.cpu().numpy()))
    if rank == 0:
        parts.sort(key=lambda t: t[0])
        sharded = np.concatenate([q[1] for q in parts], axis=1)
        single, _ = m.eval(torch.from_numpy(mesh.P).cuda())
        np.savez(a.out, sharded=sharded, single=single.cpu().numpy(), world=world)
    dist.barrier()
    m.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
