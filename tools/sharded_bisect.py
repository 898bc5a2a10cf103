#!/usr/bin/env python
"""Diagnostic (N ranks, torchrun): one particle-sharded beam against the single-GPU run of the same beam; when the loss
flags differ, bisects the end element for the first stretch on which they do and prints the particles involved.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      tools/sharded_bisect.py [n] [first] [last]
"""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from hpsim_b200 import HPSim  # noqa: E402

HPSim.print_timing = False


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    first = sys.argv[2] if len(sys.argv) > 2 else "TBDB02"
    last = sys.argv[3] if len(sys.argv) > 3 else "01QM10"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    HPSim.set_gpu(local)
    lib = H.cuda_lib()
    ids = [HPSim.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    mode = HPSim.init_sharding(rank, world, ids[0])
    case = H.Case(lib, n=n, sc=True)
    one = None
    if rank == 0:
        HPSim._ctx_cache.clear()  # a context without a communicator for the single-GPU run
        one = H.Case(lib, n=n, sc=True)
    names = case.beamline.get_element_names(first, last)

    def probe(end):
        case.beam.restore_initial_beam()
        case.sim.reset()
        case.run(first, end)
        st = case.state()
        ms = case.beam.moment_stats()
        parts = [None] * world
        dist.all_gather_object(parts, (case.beam.offset, case.beam.stride, {k: st[k] for k in ("phi", "w", "loss", "lloss")},
                                       {k: ms[k] for k in ("fused", "fallback", "deferred", "fallback_cause")}))
        out = [None]
        if rank == 0:
            full = {k: np.empty(n, parts[0][2][k].dtype) for k in parts[0][2]}
            for off, stride, part, _ in parts:
                for k in part:
                    full[k][off::stride] = part[k]
            one.beam.restore_initial_beam()
            one.sim.reset()
            one.run(first, end)
            ref = one.state()
            d = (int((full["lloss"] != ref["lloss"]).sum()), int((full["loss"] != ref["loss"]).sum()))
            good = (ref["loss"] == 0) & (ref["lloss"] == 0)
            dphi = float(np.abs(full["phi"][good] - ref["phi"][good]).max()) if good.any() else 0.0
            out = [(d, dphi, [p[3] for p in parts][:2], one.beam.moment_stats()["fallback_cause"])]
            probe.last = (full, ref)
        dist.broadcast_object_list(out, src=0)
        return out[0]

    d, dphi, ms, ms1 = probe(names[-1])
    if rank == 0:
        print("SHARDED_BISECT", json.dumps({"mode": mode, "world": world, "n": n, "range": [first, last], "flag_diff": d, "max_dphi_good": dphi,
                                            "moment_stats_rank01": ms, "single_fallback_cause": ms1}))
    if d[0] == 0 and d[1] == 0:
        return
    lo, hi = 0, len(names) - 1
    while lo < hi:
        mid = (lo + hi) // 2
        d, dphi, ms, ms1 = probe(names[mid])
        if rank == 0:
            print("  probe", mid, names[mid], d, dphi, ms[0], ms1)
        if d[0] or d[1]:
            hi = mid
        else:
            lo = mid + 1
    d, dphi, ms, ms1 = probe(names[lo])
    if rank == 0:
        full, ref = probe.last
        idx = np.nonzero(full["lloss"] != ref["lloss"])[0]
        good = (ref["loss"] == 0) & (ref["lloss"] == 0)
        print("first differing end element:", lo, names[lo], "particles", len(idx), idx[:12], "ranks", (idx[:12] % world))
        print("  phi sharded", full["phi"][idx[:8]], "single", ref["phi"][idx[:8]])
        print("  lloss sharded", full["lloss"][idx[:8]], "single", ref["lloss"][idx[:8]], "loss", full["loss"][idx[:8]], ref["loss"][idx[:8]])
        print("  <phi>_good single", ref["phi"][good].mean(), "+3pi", ref["phi"][good].mean() + 3 * np.pi, "moment stats", ms, ms1)
    if lo > 0:
        d0 = probe(names[lo - 1])
        if rank == 0:
            full, ref = probe.last
            print("  one element earlier (%s): flag diff %s max dphi good %g; phi of those particles sharded %s single %s" %
                  (names[lo - 1], d0[0], d0[1], full["phi"][idx[:8]], ref["phi"][idx[:8]]))


if __name__ == "__main__":
    main()
