"""CPU, N ranks over torch.distributed/gloo: the particle-sharded CPU oracle (interleaved ownership, every
sum / count / max / charge mesh folded over the ranks) must reproduce the single-process oracle.  This
pins the host-side logic of the multi-GPU path -- which quantities cross ranks and where in the kick
chain -- without a GPU.  Launched by tests/test_sharded_oracle.py; rank 0 prints one JSON line."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402
from hpsim_b200 import HPSim  # noqa: E402

HPSim.print_timing = False
ALLREDUCE = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)


def main():
    n, first, last, sc = int(sys.argv[1]), sys.argv[2], sys.argv[3], sys.argv[4] == "1"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = H.oracle_lib()
    lib.hps_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
    ncalls = [0]

    @ALLREDUCE
    def allreduce(buf, count, op, _user):
        a = np.ctypeslib.as_array(buf, shape=(count,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
        ncalls[0] += 1

    HPSim._use_library(lib)
    ctx = HPSim._ctx()
    lib.hps_oracle_set_threads(ctx.h, 2)
    lib.hps_oracle_set_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, ALLREDUCE, C.c_void_p]
    ctx.check(lib.hps_oracle_set_comm(ctx.h, rank, world, allreduce, None))
    case = H.Case(lib, n=n, sc=sc)
    assert case.beam.stride == world and case.beam.offset == rank
    run = case.run(first, last)
    st = case.state()
    mom = case.moments()
    calls_sharded = ncalls[0]
    parts = [None] * world
    dist.all_gather_object(parts, (case.beam.offset, case.beam.stride, st))
    dist.barrier()
    if rank != 0:
        return
    full = {k: np.empty(n, st[k].dtype) for k in st}
    for off, stride, part in parts:
        for k in st:
            full[k][off::stride] = part[k]
    # single-process oracle on a fresh, unsharded context
    HPSim._ctx_cache.clear()
    one = H.Case(lib, n=n, sc=sc)
    run1 = one.run(first, last)
    ref = one.state()
    mom1 = one.moments()
    rep = H.compare(full, ref, 1e-9)
    rep.update(world=world, n=n, kicks=(run["kicks"], run1["kicks"]), remesh=(run["remesh"], run1["remesh"]),
               allreduce_calls=calls_sharded,
               moments_rel={k: float(abs(mom[k] - mom1[k]) / max(abs(mom1[k]), 1e-300)) for k in mom1 if k not in ("loss", "lloss")},
               counts=((int(mom["loss"]), int(mom["lloss"])), (int(mom1["loss"]), int(mom1["lloss"]))))
    print("SHARDED_ORACLE " + json.dumps(rep, default=float))


if __name__ == "__main__":
    main()
