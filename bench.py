#!/usr/bin/env python
"""bench.py -- particle-element steps/s of the HPSim tracking + SCHEFF hot path on B200.

Workload (BASELINE.json configs[1], "C2"): DTL tanks 1-4 (01QM00U -> 04QM30, model indices
130..478, 349 elements, 406 space-charge kicks), 1 048 576 particles per GPU, SCHEFF 32x128 r-z
mesh, interval 0.025 m, adj_bunch 3.  The input beam is the C1-recipe DC beam (seed 20161201)
tracked TBDB02 -> spch_comp_7 once, untimed, and kept as the "intermediate" beam.

One step = restore the intermediate beam + one simulate(01QM00U, 04QM30).
  value : steps/s with the beam resident in HBM (restore is a device-to-device copy)
  e2e   : the same step through the C ABI with HOST buffers: upload of the 8 particle arrays
          from pinned memory, simulate, download of the 8 arrays -- all inside the timed region
  roofline : the fused segment kernel (one sweep = 112 B/particle algorithmic) timed with CUDA
          events between consecutive launches in one extra profiled step
  cpu_baseline : the CPU oracle (oracle/, OpenMP, all host cores) on a bounded sample

N > 1 (torchrun): weak scaling -- ONE beam of N x 1 Mi particles, particle-sharded over the N GPUs
(interleaved ownership: global particle i lives on rank i mod N); every kick exchanges three small moment
vectors and the 32 KB charge mesh between the ranks inside the reducing kernels (peer stores over NVLink;
NCCL all-gather as fallback).
--replicas instead tracks one independent 1 Mi-particle beam per GPU (scan style, BASELINE configs[4]).
`--impl reference` runs the reference's own CUDA code (oracle/_ref/ref_driver, sm_100 rebuild of
the unmodified sources) on the same workload; the reference has no CPU implementation.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

PRE = ("TBDB02", "spch_comp_7")
RANGE = ("01QM00U", "04QM30")
N_PER_GPU = 1 << 20
METRIC = "particle_element_steps_per_sec"
UNIT = "particle-element steps/s"
SEG_BYTES = 112  # per particle per fused sweep: 6 fp64 + 2 u32 read and written
# DRAM traffic of one k_segment launch at 1 Mi particles from the `ncu --set full` capture committed as
# profiles/r1b_ncu_k_segment.txt (dram__bytes_read.sum + dram__bytes_write.sum; ncu flushes the L2 before every
# replay, so this is the cold-cache figure: the sweep re-reads the whole 59 MB beam, its writes stay in L2)
SEG_TRAFFIC_1MI = 58888704 + 486656


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        smax = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    import harness as H
    exe = H.REF_SP if args.ref_float else H.REF_DP
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver_dp was not built (reference sources absent at build time)"}))
        return
    n = N_PER_GPU
    import tempfile
    work = tempfile.mkdtemp(prefix="hpsim_bench_ref_")  # the 59 MB beam dump must not land in gpurun_out/
    reps = args.steps + args.warmup
    cfg = [f"lattice {H.LATTICE}", f"n {n}", f"mass {H.MASS}", f"charge {H.CHARGE}", f"current {H.CURRENT}", "sc 1",
           "nr 32", "nz 128", "interval 0.025", "adj_bunch 3", "adj_bunch_cutoff_w 0.8", "remesh_threshold 0.02",
           "gen dc", f"seed {H.SEED}", f"freq {H.FREQ}", f"ref_w {H.REF_W}", f"ref_phi {H.PHI_OFFSET!r}", f"shift_phi {H.PHI_OFFSET!r}",
           f"pre_first {PRE[0]}", f"pre_last {PRE[1]}", f"first {RANGE[0]}", f"last {RANGE[1]}", f"repeats {reps}",
           f"beam_out {os.path.join(work, 'ref_out.bin')}"]
    for k, v in H.DC.items():
        cfg.append(f"{k} {v!r}")
    path = os.path.join(work, "ref.cfg")
    open(path, "w").write("\n".join(cfg) + "\n")
    p = subprocess.run([exe, path], capture_output=True, text=True, timeout=3000)
    if p.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": f"ref_driver failed rc={p.returncode}: {p.stderr[-200:]}"}))
        return
    walls = [float(m.group(1)) for m in re.finditer(r"ref_driver wall: ([0-9.eE+-]+) \[sec\]", p.stdout)]
    walls = walls[args.warmup:]
    lat = H.lattice()
    from hpsim_b200 import lattice as L
    nel = L.index_of(lat, RANGE[1]) - L.index_of(lat, RANGE[0]) + 1
    t = sum(walls)
    value = n * nel * len(walls) / t
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": len(walls), "warmup": args.warmup,
            "ms_per_step": 1e3 * t / len(walls), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "C2: DTL tanks 1-4 (01QM00U->04QM30), 1Mi particles, SCHEFF 32x128, adj_bunch 3",
                       "note": "reference CUDA kernels rebuilt unmodified for sm_100 (%s SCHEFF tables), run on the same B200; "
                               "the reference has no CPU implementation of this path" % ("float" if args.ref_float else "double")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                             "sample": f"{len(walls)} x simulate({RANGE[0]},{RANGE[1]}) at N={n}, reference's own wall clock around Simulate()"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline(nel_range, budget_s=20.0):
    """CPU oracle (port) on the host cores, same element range, bounded particle sample."""
    import harness as H
    from hpsim_b200 import HPSim
    HPSim.print_timing = False
    lib = H.oracle_lib()
    n = 524288
    c = H.Case(lib, n=n, sc=True)
    c.run(*PRE)
    c.beam.save_intermediate_beam()
    t0 = time.time()
    c.beam.restore_intermediate_beam()
    c.run(*RANGE)
    dt = time.time() - t0
    return {"value": n * nel_range / dt, "unit": UNIT, "cores": int(lib.hps_oracle_get_threads()), "kind": "port",
            "sample": f"1 x simulate({RANGE[0]},{RANGE[1]}) at N={n} (CPU oracle, OpenMP), {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--particles", type=int, default=N_PER_GPU)
    ap.add_argument("--replicas", action="store_true", help="N>1: one independent beam per GPU (scan style) instead of one sharded beam")
    ap.add_argument("--ref-float", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fast", action="store_true", help="skip the informational HPS_MATH_FAST timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, local, world = dist_env()
    import harness as H
    from hpsim_b200 import HPSim, _capi as K, lattice as L

    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of timings
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    HPSim.set_gpu(local)
    HPSim.print_timing = False
    lib = K.load()
    n = args.particles
    sharded = world > 1 and not args.replicas
    xr_mode = "single"
    if sharded:  # one beam of world x n particles, particle-sharded; per-kick exchanges over NVLink
        ids = [HPSim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        xr_mode = HPSim.init_sharding(rank, world, ids[0])
    case = H.Case(lib, n=n * world if sharded else n, sc=True)
    ctx = case.beam._c
    nel = case.beamline.index(RANGE[1]) - case.beamline.index(RANGE[0]) + 1

    # untimed lead-in: LEBT with space charge, then keep the DTL-entrance beam
    case.run(*PRE)
    case.beam.save_intermediate_beam()
    host = case.state()  # DTL-entrance beam on the host (for the e2e leg)
    nbytes = sum(a.nbytes for a in host.values())
    pinned = {}
    for k, a in host.items():
        p = lib.hps_host_alloc(a.nbytes)
        buf = (C.c_char * a.nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=a.dtype)
        arr[:] = a
        pinned[k] = (p, arr)
    design_w = case.beam._c.lib.hps_beam_get_ref_w(case.beam.h)  # not used; keeps the getter exercised
    order = ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")

    def barrier():
        lib.hps_ctx_synchronize(ctx.h)
        if dist is not None:
            dist.barrier()

    def step_resident():
        case.beam.restore_intermediate_beam()
        rs = K.RunStats()
        ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, None))
        return rs

    def step_e2e():
        case.beam.restore_intermediate_beam()  # restores design_w / freq / current host members
        ctx.check(lib.hps_beam_upload(case.beam.h, *[pinned[k][0] for k in order]))
        ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, None))
        ctx.check(lib.hps_beam_download(case.beam.h, *[pinned[k][0] for k in order]))

    first, last = case.beamline.index(RANGE[0]), case.beamline.index(RANGE[1])

    def timed(step_fn, steps):
        """Per-step CUDA-event time on the context's stream; L2 flushed (384 MB memset) before each
        step, outside the event pair; the step's own host-side submission overlaps its execution."""
        total = 0.0
        for _ in range(steps):
            ctx.check(lib.hps_ctx_flush_l2(ctx.h))
            barrier()
            ctx.check(lib.hps_ctx_event_record(ctx.h, 0))
            step_fn()
            ctx.check(lib.hps_ctx_event_record(ctx.h, 1))
            ms = C.c_double()
            ctx.check(lib.hps_ctx_event_elapsed(ctx.h, 0, 1, C.byref(ms)))
            total += ms.value * 1e-3
        return total

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = lib.hps_ctx_launch_count(ctx.h)
    sampler = ClockSampler(local)
    sampler.start()
    dt = timed(step_resident, args.steps)
    launches = lib.hps_ctx_launch_count(ctx.h) - l0
    # ---- end-to-end timing (host buffers; upload + simulate + download inside the event pair)
    te = None
    if not args.no_e2e:
        e2e_in = {k: pinned[k][1].copy() for k in order}

        def e2e_once():
            step_e2e()

        for _ in range(max(1, args.warmup // 2)):
            for k in order:  # the download of the previous step must not become this step's input
                pinned[k][1][:] = e2e_in[k]
            step_e2e()
        te = 0.0
        for _ in range(args.steps):
            for k in order:
                pinned[k][1][:] = e2e_in[k]
            te += timed(e2e_once, 1)
    clocks = sampler.summary()
    # ---- informational: the same resident step with the opt-in fast maps (hps_ctx_set_math)
    tf = None
    math_mode = HPSim.get_math()
    if not args.no_fast and math_mode != "fast":
        HPSim.set_math("fast")
        step_resident()
        tf = timed(step_resident, max(2, args.steps // 2)) / max(2, args.steps // 2)
        HPSim.set_math(math_mode)

    # ---- one profiled step for the roofline of the dominant kernel
    lib.hps_ctx_set_profile(ctx.h, 1)
    case.beam.restore_intermediate_beam()
    rs = K.RunStats()
    ph0, ph1 = (C.c_uint64 * 8)(), (C.c_uint64 * 8)()
    lib.hps_scheff_phase_cycles(case.spch.h, ph0)
    ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, C.byref(rs)))
    lib.hps_scheff_phase_cycles(case.spch.h, ph1)
    ms = (C.c_double * 5)()
    cnt = (C.c_uint64 * 5)()
    lib.hps_ctx_get_profile(ctx.h, ms, cnt)
    lib.hps_ctx_set_profile(ctx.h, 0)
    prof = {K.PROF_CLASSES[i]: {"ms": ms[i], "launches": int(cnt[i])} for i in range(5)}
    if prof["deposit"]["launches"] == 0 and prof["moments"]["launches"] > 0:
        # fused chain: moments + deposit + field solve are ONE launch (k_scheff), timed under the "moments" class
        prof["scheff(moments+deposit+field)"] = prof.pop("moments")
        for k in ("lloss", "deposit", "field"):
            prof.pop(k)

    if dist is not None:
        import torch
        t = torch.tensor([dt, te if te is not None else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, te = float(t[0]), (float(t[1]) if te is not None else None)
    if rank != 0:
        return
    total_particles = n * world
    value = total_particles * nel * args.steps / dt
    e2e = total_particles * nel * args.steps / te if te else None
    peak, peak_kind = peaks()
    seg = prof["segment"]
    seg_us = 1e3 * seg["ms"] / max(1, seg["launches"])
    achieved = SEG_BYTES * n / (seg_us * 1e-6) / 1e9
    run = rs.as_dict()
    bytes_alg = n * (168 * run["kicks"] + 32 * run["remesh"] + 112 * run["extra_sweeps"])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: DTL tanks 1-4 (01QM00U->04QM30, 349 elements, %d SC kicks), %d particles/GPU, SCHEFF 32x128, "
                               "interval 0.025 m, adj_bunch 3; input = C1 DC beam tracked TBDB02->spch_comp_7 (untimed)" % (run["kicks"], n),
                   "math": math_mode + " (the reference's expression order, IEEE-exact division / sqrt)" if math_mode == "reference" else math_mode,
                   "parallelism": ("single GPU" if world == 1 else
                                   "one beam of %d particles sharded over %d GPUs (interleaved ownership); per kick 3 moment "
                                   "exchanges + the 32 KB r-z mesh, mode=%s" % (n * world, world, xr_mode) if sharded else
                                   "1 independent beam per GPU (replicas, no communication)"),
                   "l2": "L2 flushed (384 MB memset) before every timed step; inside a step the %.0f MB beam is swept ~%d times "
                         "and largely stays in the 126 MB L2 -- that is the workload" % (56 * n / 1e6, run["sweeps"] + 2 * run["kicks"]),
                   "timing": "CUDA events on the context's stream around each step, summed over steps; max over ranks",
                   "run": run},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": 1e3 * te / args.steps if te else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_segment", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": SEG_TRAFFIC_1MI if n == N_PER_GPU else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, cold L2 (profiles/r1b_ncu_k_segment.txt)",
                     "peak_kind": peak_kind,
                     "avg_launch_us": seg_us, "launches_per_step": seg["launches"],
                     "whole_step": {"bytes_alg": bytes_alg, "achieved_gbs": bytes_alg / (1e-3 * run["device_ms"]) / 1e9,
                                    "frac": bytes_alg / (1e-3 * run["device_ms"]) / 1e9 / peak},
                     "by_kernel_ms": {k: v["ms"] for k, v in prof.items()}},
    }
    if any(ph1[i] != ph0[i] for i in range(6)):
        # block 0's SM-clock cycles per phase of the fused chain in the profiled step, as ms at the sampled SM clock
        mhz = clocks.get("sm_mhz") or 1965.0
        names = ("moment_sweep", "reduce_policy_barrier", "deposit_flush", "green_slice", "barrier", "field_solve")
        line["roofline"]["scheff_phase_ms"] = {nm: (ph1[i] - ph0[i]) / (mhz * 1e3) for i, nm in enumerate(names)}
    if tf:
        line["fast_math"] = {"ms_per_step": 1e3 * tf, "value": total_particles * nel / tf, "unit": UNIT,
                             "note": "informational: hps_ctx_set_math(HPS_MATH_FAST), device-resident step; not the benched default"}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(nel)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
