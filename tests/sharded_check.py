"""GPU box, N ranks (torchrun): one particle-sharded beam over N GPUs must reproduce the single-GPU
run of the same beam -- loss flags bit-exact, coordinates to 1e-9 of an rms (the only difference is the
summation order of the moment / mesh reductions).  Rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/sharded_check.py [n] [first] [last] [mesh_size_cutoff_w]
"""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402
from hpsim_b200 import HPSim  # noqa: E402

HPSim.print_timing = False


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    first = sys.argv[2] if len(sys.argv) > 2 else "TBDB02"
    last = sys.argv[3] if len(sys.argv) > 3 else "01QM10"
    mesh_w = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0   # optional: SetMeshSizeCutoffW
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    HPSim.set_gpu(local)
    lib = H.cuda_lib()
    ids = [HPSim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    mode = HPSim.init_sharding(rank, world, ids[0])

    case = H.Case(lib, n=n, sc=True)
    if mesh_w:
        case.spch.set_mesh_size_cutoff_w(mesh_w)
    run = case.run(first, last)
    st = case.state()            # same call sequence as the single-GPU run below: every Update* call can
    mom = case.moments()         # mark more longitudinal losses (H3), so the order matters
    nloss, nlloss = int(mom["loss"]), int(mom["lloss"])
    parts = [None] * world
    dist.all_gather_object(parts, (case.beam.offset, case.beam.stride, st))
    if rank != 0:
        return
    full = {k: np.empty(n, st[k].dtype) for k in st}
    for off, stride, part in parts:  # local particle j of a rank is global particle off + j * stride
        for k in st:
            full[k][off::stride] = part[k]
    # single-GPU run of the same beam on a context without a communicator
    HPSim._ctx_cache.clear()  # the sharded objects keep their own context alive
    one = H.Case(lib, n=n, sc=True)
    if mesh_w:
        one.spch.set_mesh_size_cutoff_w(mesh_w)
    run1 = one.run(first, last)
    ref = one.state()
    rep = H.compare(full, ref, 1e-9)
    rep.update(mode=mode, world=world, n=n, range=[first, last], kicks=run["kicks"], remesh=(run["remesh"], run1["remesh"]),
               n_loss_sharded=nloss, n_lloss_sharded=nlloss, device_ms=(run["device_ms"], run1["device_ms"]),
               moments_sharded={k: float(v) for k, v in mom.items()},
               moments_single={k: float(v) for k, v in one.moments().items()})
    rep["ok"] = bool(rep["ok"] and rep["lloss_equal"] and run["remesh"] == run1["remesh"])
    print("SHARDED " + json.dumps(rep))
    if not rep["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
