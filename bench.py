#!/usr/bin/env python
"""bench.py -- particle-element steps/s of the HPSim tracking + SCHEFF hot path on B200.

Default workload = the configuration BASELINE.json's metric is quoted on ("C4", configs[3]): the full LANSCE
linac TBDB02 -> 48DT (model indices 62..6537: 6 476 elements, 8 894 space-charge kicks, LEBT + DTL + the
201.25 -> 805 MHz jump + all CCL modules), ONE beam of 16 Mi particles (C1 recipe, pytest/sim-lbeg.py:59-103),
SCHEFF 32x128, interval 0.025 m, adj_bunch 3.  `--gpus N` shards that one beam over N GPUs (16 Mi / N particles
each, interleaved ownership): STRONG scaling.  Other BASELINE configurations: `--config c1|c2|c3|c5`.

One step = restore the saved input beam (device-to-device) + one simulate(first, last).
  value : steps/s with the beam resident in HBM
  e2e   : the same step through the C ABI with HOST buffers: upload of the 8 particle arrays from pinned
          memory, simulate, download of the 8 arrays -- all inside the timed region (fewer steps than --steps
          when a step takes seconds; the count is stated)
  roofline : the fused segment kernel (one sweep = 112 B/particle algorithmic) timed with CUDA events between
          consecutive launches in one extra profiled step; whole-step fraction from N(168 K + 32 R + 112 P)
  cpu_baseline : the CPU oracle (oracle/, OpenMP, all host cores) on a bounded sample of the same range
N > 1: per kick the ranks exchange the moment vectors and the 32 KB charge mesh inside the reducing kernels (peer
stores over NVLink; NCCL all-gather as fallback); rank 0 re-runs the beam on one GPU afterwards and the loss
counts / energy sum of the sharded run are checked against it (`sharded_check`).
`--impl reference` times the reference's own CUDA code (oracle/_ref/ref_driver*, sm_100 rebuild of the unmodified
sources; the reference has no CPU implementation) on the SAME element range with a bounded particle count.
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

METRIC = "particle_element_steps_per_sec"
UNIT = "particle-element steps/s"
SEG_BYTES = 112  # per particle per fused sweep: 6 fp64 + 2 u32 read and written

# name -> workload.  n_total: one beam split over the GPUs (strong); n_per_gpu: work per GPU fixed (weak)
CONFIGS = {
    "c4": dict(label="C4: full LANSCE linac TBDB02->48DT", pre=None, rng=("TBDB02", "48DT"), n_total=1 << 24, remesh=0.02),
    "c3": dict(label="C3: LEBT+DTL+CCL module 5, TBDB02->05DT", pre=None, rng=("TBDB02", "05DT"), n_total=1 << 22, remesh=0.02),
    "c2": dict(label="C2: DTL tanks 1-4, 01QM00U->04QM30 (input = C1 beam tracked TBDB02->spch_comp_7, untimed)",
               pre=("TBDB02", "spch_comp_7"), rng=("01QM00U", "04QM30"), n_per_gpu=1 << 20, remesh=0.02),
    "c1": dict(label="C1: the reference script's own workload (pytest/sim-lbeg.py), TBDB02->48DT, 64 Ki particles",
               pre=None, rng=("TBDB02", "48DT"), n_total=1 << 16, remesh=0.02),
    "c5": dict(label="C5: buncher set-point scan (sim-lbeg-scan-pv.py style), TBDB02->TREM01, one independent beam per GPU",
               pre=None, rng=("TBDB02", "TREM01"), n_per_gpu=1 << 20, remesh=0.2, replicas=True),
}
# measured DRAM traffic of one k_segment launch (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, cold L2):
# local particles -> (bytes, source)
SEG_TRAFFIC = {1 << 20: (58888704 + 486656, "profiles/r1b_ncu_k_segment.txt")}
_t = os.path.join(ROOT, "profiles", "r2_seg_traffic.json")
if os.path.exists(_t):
    for _k, _v in json.load(open(_t)).items():
        SEG_TRAFFIC[int(_k)] = (_v["bytes"], _v["source"])
# what bounds the sweep when it is not HBM (ncu --set full of the dominant pattern sweep: fp64 pipe / issue-slot utilisation)
SEG_COMPUTE = None
_t = os.path.join(ROOT, "profiles", "r2_seg_compute.json")
if os.path.exists(_t):
    SEG_COMPUTE = json.load(open(_t))
# particle count of the reference arm.  Measured on B200 (profiles/r2_reference_sweep.txt), TBDB02->48DT, double tables:
# 64 Ki particles 5.1 s per pass (8.3e7 steps/s), 256 Ki 11.7 s (1.45e8 steps/s): the reference is launch-bound below and
# atomics-bound above; 25 passes (--steps 20 --warmup 5) at 256 Ki take ~5 min, 1 Mi would take ~15
REF_N = {"c4": 1 << 18, "c1": 1 << 16, "c3": 1 << 19, "c2": 1 << 20, "c5": 1 << 19}


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
            time.sleep(0.25)

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


def workload_text(cfg, n_total, world, kicks=None):
    return ("%s, %d particles in total%s, SCHEFF 32x128, interval 0.025 m, adj_bunch 3 (cutoff 0.8 MeV), remesh threshold %g%s"
            % (cfg["label"], n_total, " (%d per GPU)" % (n_total // world) if world > 1 else "", cfg["remesh"],
               ", %d SC kicks" % kicks if kicks else ""))


# ------------------------------------------------------------------------------- reference arm
def ref_pass(exe, cfgname, n, reps, work):
    """`reps` x simulate(range) by the reference's own code; returns the per-pass wall times it measured itself."""
    import harness as H
    cfg = CONFIGS[cfgname]
    lines = [f"lattice {H.LATTICE}", f"n {n}", f"mass {H.MASS}", f"charge {H.CHARGE}", f"current {H.CURRENT}", "sc 1",
             "nr 32", "nz 128", "interval 0.025", "adj_bunch 3", "adj_bunch_cutoff_w 0.8", f"remesh_threshold {cfg['remesh']}",
             "gen dc", f"seed {H.SEED}", f"freq {H.FREQ}", f"ref_w {H.REF_W}", f"ref_phi {H.PHI_OFFSET!r}", f"shift_phi {H.PHI_OFFSET!r}",
             f"first {cfg['rng'][0]}", f"last {cfg['rng'][1]}", f"repeats {reps}", f"beam_out {os.path.join(work, 'ref_out.bin')}"]
    if cfg["pre"]:
        lines += [f"pre_first {cfg['pre'][0]}", f"pre_last {cfg['pre'][1]}"]
    for k, v in H.DC.items():
        lines.append(f"{k} {v!r}")
    path = os.path.join(work, "ref.cfg")
    open(path, "w").write("\n".join(lines) + "\n")
    p = subprocess.run([exe, path], capture_output=True, text=True, timeout=3000)
    if p.returncode != 0:
        raise RuntimeError(f"ref_driver failed rc={p.returncode}: {p.stderr[-200:]}")
    return [float(m.group(1)) for m in re.finditer(r"ref_driver wall: ([0-9.eE+-]+) \[sec\]", p.stdout)]


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    import harness as H
    import tempfile
    from hpsim_b200 import lattice as L
    if not os.path.exists(H.REF_DP):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver_dp was not built (reference sources absent at build time)"}))
        return
    cfg = CONFIGS[args.config]
    n = args.ref_particles or REF_N[args.config]
    lat = H.lattice()
    nel = L.index_of(lat, cfg["rng"][1]) - L.index_of(lat, cfg["rng"][0]) + 1
    work = tempfile.mkdtemp(prefix="hpsim_bench_ref_")  # beam dumps must not land in gpurun_out/
    try:
        walls = ref_pass(H.REF_SP if args.ref_float else H.REF_DP, args.config, n, args.steps + args.warmup, work)[args.warmup:]
        other = None
        if not args.ref_float and os.path.exists(H.REF_SP):  # the shipped default build (float SCHEFF tables), short
            w2 = ref_pass(H.REF_SP, args.config, n, 3, work)[1:]
            other = {"value": n * nel * len(w2) / sum(w2), "unit": UNIT, "ms_per_step": 1e3 * sum(w2) / len(w2), "steps": len(w2),
                     "note": "the reference's default build (make_root.mk:22-23: float SCHEFF tables, float atomics), same range and N"}
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:300]}))
        return
    t = sum(walls)
    value = n * nel * len(walls) / t
    n_ours = cfg.get("n_total") or cfg["n_per_gpu"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": len(walls), "warmup": args.warmup,
            "ms_per_step": 1e3 * t / len(walls), "higher_is_better": True, "scaling": "strong" if "n_total" in cfg else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_text(cfg, n_ours, 1),
                       "reference_particles": n,
                       "note": "reference CUDA kernels rebuilt unmodified for sm_100 (%s SCHEFF tables), same B200, same element range "
                               "and SCHEFF settings, N=%d particles instead of %d so that --steps + --warmup passes end within minutes; "
                               "steps/s is throughput-normalised (particles x elements / time), so the ratio to the other arm compares "
                               "throughputs, not equal-size runs.  The reference has no CPU implementation of this path."
                               % ("float" if args.ref_float else "double", n, n_ours)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                             "sample": f"{len(walls)} x simulate({cfg['rng'][0]},{cfg['rng'][1]}) at N={n}, reference's own wall clock around Simulate()"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if other:
        line["reference_float_build"] = other
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline(cfg, nel, n):
    """CPU oracle (port) on the host cores, same element range, bounded particle sample."""
    import harness as H
    from hpsim_b200 import HPSim
    HPSim.print_timing = False
    lib = H.oracle_lib()
    c = H.Case(lib, n=n, sc=True, remesh=cfg["remesh"])
    if cfg["pre"]:
        c.run(*cfg["pre"])
    t0 = time.time()
    c.run(*cfg["rng"])
    dt = time.time() - t0
    return {"value": n * nel / dt, "unit": UNIT, "cores": int(lib.hps_oracle_get_threads()), "kind": "port",
            "sample": f"1 x simulate({cfg['rng'][0]},{cfg['rng'][1]}) at N={n} (CPU oracle, OpenMP), {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--particles", type=int, default=0, help="override the particle count (total for c1/c3/c4, per GPU for c2/c5)")
    ap.add_argument("--replicas", action="store_true", help="N>1: one independent beam per GPU (scan style) instead of one sharded beam")
    ap.add_argument("--ref-float", action="store_true", help="--impl reference: time the float-table build as the main line")
    ap.add_argument("--ref-particles", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: min(--steps, 3) when a step takes > 1 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the single-GPU re-run that checks the sharded result")
    ap.add_argument("--fast", action="store_true", help="add the informational HPS_MATH_FAST timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, local, world = dist_env()
    import harness as H
    from hpsim_b200 import HPSim, _capi as K

    cfg = CONFIGS[args.config]
    replicas = args.replicas or cfg.get("replicas", False)
    strong = "n_total" in cfg and not replicas
    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of timings
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    HPSim.set_gpu(local)
    HPSim.print_timing = False
    lib = K.load()
    if strong:
        n_total = args.particles or cfg["n_total"]
        n_beam = n_total                      # one beam, sharded when world > 1
    else:
        n_per = args.particles or cfg.get("n_per_gpu") or cfg["n_total"]
        n_total = n_per * world
        n_beam = n_per if replicas else n_total
    sharded = world > 1 and not replicas
    n_local = n_beam // world if sharded else n_beam
    xr_mode = "single"
    if sharded:  # one beam particle-sharded over the ranks; per-kick exchanges over NVLink
        ids = [HPSim.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        xr_mode = HPSim.init_sharding(rank, world, ids[0])
    case = H.Case(lib, n=n_beam, sc=True, remesh=cfg["remesh"])
    ctx = case.beam._c
    first, last = case.beamline.index(cfg["rng"][0]), case.beamline.index(cfg["rng"][1])
    nel = last - first + 1
    if cfg.get("replicas"):
        # C5: every replica tracks its own main-buncher amplitude, spread over +-5 %; the PV feeds only
        # the voltage of element TDDB01 (buncher p[0]), pushed through hps_lattice_update_elem like set_db_epics does
        from hpsim_b200 import lattice as L
        tab = case.beamline.table.copy()
        i = L.index_of(tab, "TDDB01")
        tab["p"][i][0] *= 1.0 + 0.05 * (2.0 * (rank + 0.5) / world - 1.0)
        case.beamline._refresh(tab)

    # ---- sharded result against a single-GPU run of the same beam (rank 0, its own context without a communicator).
    # (1) EXACT on a short range, both beams fresh (LEBT + the first DTL cells, 258 kicks: every exchange of the kick chain
    #     is exercised hundreds of times and the dynamics have not yet amplified the summation-order difference): loss
    #     counts equal, sums of w / phi / x^2 over the survivors to 1e-9.  It has to come first: the reference's stale
    #     counters make a pass depend on the passes before it.
    # (2) REPORTED on the timed range after the timing: through the whole linac the space-charge dynamics are chaotic (two
    #     code versions differ by 2 % in the longitudinal-loss count), so the counts of the timed step are compared loosely.
    check = None
    if sharded and not args.no_check:
        def sums(st):
            a = st["loss"] == 0
            return np.array([float((st["loss"] != 0).sum()), float((st["lloss"] != 0).sum()), float(st["w"][a].sum()),
                             float(st["phi"][a].sum()), float((st["x"][a] ** 2).sum()), float(a.sum())])
        short = ("TBDB02", "01QM10")
        case.run(*short)
        import torch
        t = torch.tensor(sums(case.state()), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ss = t.numpy()
        case.beam.restore_initial_beam()
        if rank == 0:
            keep = dict(HPSim._ctx_cache)
            HPSim._ctx_cache.clear()
            one = H.Case(lib, n=n_beam, sc=True, remesh=cfg["remesh"])
            one.run(*short)
            s1 = sums(one.state())
            del one
            HPSim._ctx_cache.clear()
            HPSim._ctx_cache.update(keep)
            exact_ok = bool(ss[0] == s1[0] and ss[1] == s1[1] and ss[5] == s1[5] and
                            all(abs(ss[k] - s1[k]) <= 1e-9 * abs(s1[k]) for k in (2, 3, 4)))
            check = {"short_range": {"range": list(short), "n_loss": [int(ss[0]), int(s1[0])], "n_lloss": [int(ss[1]), int(s1[1])],
                                     "sum_w": [ss[2], s1[2]], "sum_phi": [ss[3], s1[3]], "sum_x2": [ss[4], s1[4]], "ok": exact_ok,
                                     "gate": "counts equal, sums within 1e-9"}, "order": "[sharded, single GPU]", "ok": exact_ok}
        dist.barrier()

    if cfg["pre"]:  # untimed lead-in; the timed range starts from the kept "intermediate" beam
        case.run(*cfg["pre"])
        case.beam.save_intermediate_beam()
        restore = case.beam.restore_intermediate_beam
    else:
        restore = case.beam.restore_initial_beam
    host = case.state()  # input beam on the host (for the e2e leg)
    nbytes = sum(a.nbytes for a in host.values())
    order = ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")
    pinned = {}
    if not args.no_e2e:
        for k, a in host.items():
            p = lib.hps_host_alloc(a.nbytes)
            buf = (C.c_char * a.nbytes).from_address(p)
            arr = np.frombuffer(buf, dtype=a.dtype)
            arr[:] = a
            pinned[k] = (p, arr)

    def barrier():
        lib.hps_ctx_synchronize(ctx.h)
        if dist is not None:
            dist.barrier()

    def step_resident():
        restore()
        ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, None))

    def step_e2e():
        restore()  # restores design_w / freq / current host members
        ctx.check(lib.hps_beam_upload(case.beam.h, *[pinned[k][0] for k in order]))
        ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, None))
        ctx.check(lib.hps_beam_download(case.beam.h, *[pinned[k][0] for k in order]))

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
    # checksum of the step's result (the last timed step): loss counts and the energy sum of the survivors
    res = case.state()
    chk_local = np.array([float((res["loss"] != 0).sum()), float((res["lloss"] != 0).sum()),
                          float(res["w"][res["loss"] == 0].sum()), float(len(res["w"]))])
    # ---- end-to-end timing (host buffers; upload + simulate + download inside the event pair)
    te, e2e_steps = None, 0
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or (args.steps if dt / args.steps < 1.0 else min(args.steps, 3))
        e2e_in = {k: pinned[k][1].copy() for k in order}
        for _ in range(1 if dt / args.steps >= 1.0 else max(1, args.warmup // 2)):
            for k in order:  # the download of the previous step must not become this step's input
                pinned[k][1][:] = e2e_in[k]
            step_e2e()
        te = 0.0
        for _ in range(e2e_steps):
            for k in order:
                pinned[k][1][:] = e2e_in[k]
            te += timed(step_e2e, 1)
        del e2e_in
    clocks = sampler.summary()
    # ---- informational: the same resident step with the opt-in fast maps (hps_ctx_set_math)
    tf = None
    math_mode = HPSim.get_math()
    if args.fast and math_mode != "fast":
        HPSim.set_math("fast")
        step_resident()
        tf = timed(step_resident, 2) / 2
        HPSim.set_math(math_mode)

    # ---- one profiled step for the roofline of the dominant kernel
    lib.hps_ctx_set_profile(ctx.h, 1)
    restore()
    rs = K.RunStats()
    ctx.check(lib.hps_simulate(ctx.h, case.beam.h, case.beamline.h, case.spch.h, first, last, C.byref(rs)))
    ms = (C.c_double * 5)()
    cnt = (C.c_uint64 * 5)()
    lib.hps_ctx_get_profile(ctx.h, ms, cnt)
    lib.hps_ctx_set_profile(ctx.h, 0)
    prof = {K.PROF_CLASSES[i]: {"ms": ms[i], "launches": int(cnt[i])} for i in range(5)}

    chk = chk_local
    if dist is not None:
        import torch
        t = torch.tensor([dt, te if te is not None else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, te = float(t[0]), (float(t[1]) if te is not None else None)
        if sharded:
            t = torch.tensor(chk_local, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            chk = t.numpy()
    # ---- the timed range, sharded vs single GPU (see the short-range check above for why this one is loose)
    if check is not None and rank == 0:
        HPSim._ctx_cache.clear()
        one = H.Case(lib, n=n_beam, sc=True, remesh=cfg["remesh"])
        if cfg["pre"]:
            one.run(*cfg["pre"])
            one.beam.save_intermediate_beam()
        for _ in range(2):  # like a timed step, the checked pass starts from a restored beam after a pass (H4 / H21)
            (one.beam.restore_intermediate_beam if cfg["pre"] else one.beam.restore_initial_beam)()
            one.run(*cfg["rng"])
        r1 = one.state()
        c1 = [float((r1["loss"] != 0).sum()), float((r1["lloss"] != 0).sum()), float(r1["w"][r1["loss"] == 0].sum())]
        loose_ok = bool(abs(chk[0] - c1[0]) <= 0.01 * n_beam and abs(chk[1] - c1[1]) <= 0.03 * n_beam and abs(chk[2] - c1[2]) <= 0.01 * abs(c1[2]))
        check["timed_range"] = {"n_loss": [int(chk[0]), int(c1[0])], "n_lloss": [int(chk[1]), int(c1[1])], "sum_w_alive": [chk[2], c1[2]],
                                "ok": loose_ok, "gate": "chaotic range: counts within 1 % / 3 % of N, energy sum within 1 %"}
        check["ok"] = bool(check["short_range"]["ok"] and loose_ok)
        del one, r1
    if dist is not None:
        dist.barrier()
    if rank != 0:
        return
    value = n_total * nel * args.steps / dt
    e2e = n_total * nel * e2e_steps / te if te else None
    peak, peak_kind = peaks()
    seg = prof["segment"]
    seg_us = 1e3 * seg["ms"] / max(1, seg["launches"])
    achieved = SEG_BYTES * n_local / (seg_us * 1e-6) / 1e9
    run = rs.as_dict()
    bytes_alg = n_local * (168 * run["kicks"] + 32 * run["remesh"] + 112 * run["extra_sweeps"])
    step_s = dt / args.steps
    traffic = SEG_TRAFFIC.get(n_local)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(cfg, n_total, world, run["kicks"]),
                   "elements": nel, "particles_total": n_total, "particles_per_gpu": n_local,
                   "math": math_mode + " (the reference's expression order, IEEE-exact division / sqrt)" if math_mode == "reference" else math_mode,
                   "parallelism": ("single GPU" if world == 1 else
                                   "one beam of %d particles sharded over %d GPUs (interleaved ownership); per kick the moment "
                                   "exchanges + the 32 KB r-z mesh, mode=%s" % (n_beam, world, xr_mode) if sharded else
                                   "1 independent beam per GPU (replicas, no communication)"),
                   "l2": "L2 flushed (384 MB memset) before every timed step; the particle state of one GPU is %.0f MB (L2: 126 MB)" % (56 * n_local / 1e6),
                   "timing": "CUDA events on the context's stream around each step, summed over steps; max over ranks",
                   "run": run},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "ms_per_step": 1e3 * te / e2e_steps if te else None, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result": {"n_loss": int(chk[0]), "n_lloss": int(chk[1]), "sum_w_alive": float(chk[2])},
        "moment_chain": case.beam.moment_stats(),
        "roofline": {"bound": "hbm", "kernel": "k_segment", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic[0] if traffic else None,
                     "traffic_source": ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, cold L2 (%s)" % traffic[1]) if traffic else None,
                     "peak_kind": peak_kind,
                     "avg_launch_us": seg_us, "launches_per_step": seg["launches"],
                     "alg_bytes_per_launch": SEG_BYTES * n_local,
                     "whole_step": {"bytes_alg": bytes_alg, "formula": "N_local (168 K + 32 R + 112 P)",
                                    "achieved_gbs": bytes_alg / step_s / 1e9, "frac": bytes_alg / step_s / 1e9 / peak,
                                    "strict_floor_frac": n_local * (112 * (run["kicks"] + run["extra_sweeps"]) + 32 * run["remesh"]) / step_s / 1e9 / peak},
                     "by_kernel_ms": {k: v["ms"] for k, v in prof.items()},
                     "by_kernel_us_per_launch": {k: 1e3 * v["ms"] / v["launches"] for k, v in prof.items() if v["launches"]},
                     "profiled_step_ms": run["device_ms"]},
    }
    if SEG_COMPUTE:  # the sweep is bound by fp64 instruction issue, not by HBM: the measured pipe utilisation says how close it is
        line["roofline"]["compute_bound_evidence"] = SEG_COMPUTE
    if sharded:
        cs = (C.c_uint64 * 4)()
        lib.hps_ctx_comm_stats(ctx.h, cs)
        line["exchange"] = {"small_exchanges": int(cs[1]), "small_wait_us_avg": cs[0] / max(1, cs[1]) / 1e3,
                            "mesh_exchanges": int(cs[3]), "mesh_wait_us_avg": cs[2] / max(1, cs[3]) / 1e3,
                            "note": "rank 0, whole process: time from a reduction's own stores being issued to every peer's words "
                                    "having arrived (NVLink latency + rank skew); mesh: block 0 of the field solve after its Green slice"}
    if check is not None:
        line["sharded_check"] = check
    if tf:
        line["fast_math"] = {"ms_per_step": 1e3 * tf, "value": n_total * nel / tf, "unit": UNIT,
                             "note": "informational: hps_ctx_set_math(HPS_MATH_FAST), device-resident step; not the benched default"}
    if not args.no_cpu:
        # bounded sample: ~10-30 s of CPU work on the same range
        line["cpu_baseline"] = cpu_baseline(cfg, nel, 65536 if nel > 1000 else 524288)
    print(json.dumps(line))
    if check is not None and not check["ok"]:
        sys.exit("sharded result differs from the single-GPU run: %s" % json.dumps(check))


if __name__ == "__main__":
    main()
