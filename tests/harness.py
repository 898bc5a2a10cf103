"""Shared helpers of the test-suite: build/load the CPU oracle, drive either implementation of the
C ABI through the HPSim mirror, run the reference-CUDA driver (oracle/_ref) and compare beams."""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hpsim_b200 import HPSim, _capi as K, lattice as L  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libhpsim_oracle.so")
REF_DP = os.path.join(ROOT, "oracle", "_ref", "ref_driver_dp")
REF_SP = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
LATTICE = os.path.join(ROOT, "tests", "golden", "lansce.lat")
# the reference's linac databases + the sqlite math extension their triggers need, vendored as test data
DB_DIR = os.path.join(ROOT, "tests", "golden", "db")
SQLITE_EXT = os.path.join(DB_DIR, "libsqliteext.so")
DB_FILES = [os.path.join(DB_DIR, f) for f in ("tbtd.db", "dtl.db", "trst.db", "ccl.db")]

# C1 beam recipe
DC = dict(ax=0.095, bx=47.0, ex=0.00327, ay=-0.102, by=60.0, ey=0.002514, dphi=np.pi, sync_phi=0.0, sync_w=0.7518)
SEED = 20161201
PHI_OFFSET = -707.907480574  # rad = -40560.11 deg
MASS, CHARGE, CURRENT, FREQ, REF_W = 939.294, -1.0, 0.015, 201.25, 0.750

_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        src = os.path.join(ROOT, "oracle", "hpsim_oracle.cpp")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "cpu"])
        _oracle = K.bind(C.CDLL(ORACLE_SO, mode=C.RTLD_LOCAL))
        _oracle.hps_oracle_get_threads.restype = C.c_int
    return _oracle


def cuda_lib():
    return K.load()


def build_tool(name):
    """Path of a helper executable under tests/_build (built by __graft_entry__.build()); compiled on demand with nvcc
    when it is missing or older than its sources."""
    import shutil
    srcs = {"gpu_math_check": ["tests/gpu_math_check.cu"],
            "gpu_maps_check": ["tests/gpu_maps_check.cu", "hpsim_b200/csrc/hps_plan.cpp"],
            "host_maps": ["tests/host_maps.cu", "hpsim_b200/csrc/hps_plan.cpp"]}[name]
    srcs = [os.path.join(ROOT, s) for s in srcs]
    deps = srcs + [os.path.join(ROOT, "hpsim_b200", "csrc", h) for h in ("hps_kernels.cuh", "hps_math.cuh", "hps_internal.h")]
    exe = os.path.join(ROOT, "tests", "_build", name)
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        flags = ["-O2", "-std=c++17", "-x", "cu", "-arch=sm_100a"] if name == "host_maps" else \
                ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"]
        subprocess.check_call([nvcc] + flags + ["-o", exe] + srcs)
    return exe


class math_mode:
    """with math_mode("reference"): ... -- arithmetic of the segment kernel on the current device (hps_ctx_set_math)."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        HPSim._use_library(cuda_lib())
        self.prev = HPSim.get_math()
        HPSim.set_math(self.mode)
        HPSim._use_library(None)

    def __exit__(self, *exc):
        HPSim._use_library(cuda_lib())
        HPSim.set_math(self.prev)
        HPSim._use_library(None)


_lat = None


def lattice():
    global _lat
    if _lat is None:
        _lat = L.load(LATTICE)
    return _lat


class Case:
    """One beam + beamline + (optional) SCHEFF on one implementation."""

    def __init__(self, lib, n=65536, sc=True, nr=32, nz=128, interval=0.025, adj_bunch=3, cutoff_w=0.8, remesh=0.02,
                 seed=SEED, table=None, make_beam=True):
        HPSim._use_library(lib)
        self.lib = lib
        self.beamline = HPSim.BeamLine(table=lattice() if table is None else table)
        self.beam = HPSim.Beam(mass=MASS, charge=CHARGE, current=CURRENT, num=n)
        if make_beam:
            self.beam.set_dc(DC["ax"], DC["bx"], DC["ex"], DC["ay"], DC["by"], DC["ey"], DC["dphi"], DC["sync_phi"], DC["sync_w"], seed)
            self.beam.set_frequency(FREQ)
            self.beam.set_ref_w(REF_W)
            self.beam.set_ref_phi(PHI_OFFSET)
            self.beam.translate("phi", PHI_OFFSET)
            self.beam.save_initial_beam()
        self.spch = None
        if sc:
            self.spch = HPSim.SpaceCharge(nr=nr, nz=nz, interval=interval, adj_bunch=adj_bunch)
            self.spch.set_adj_bunch_cutoff_w(cutoff_w)
            self.spch.set_remesh_threshold(remesh)
        self.sim = HPSim.Simulator(self.beam, self.beamline, self.spch)
        HPSim._use_library(None)

    def run(self, first, last):
        self.sim.simulate(first, last)
        return self.sim.last_run

    def state(self):
        d = self.beam._arrays()
        return d

    def moments(self):
        b = self.beam
        return dict(loss=b.get_loss_num("t"), lloss=b.get_loss_num("l"),
                    x_avg=b.get_avg_x(), y_avg=b.get_avg_y(), x_sig=b.get_sig_x(), y_sig=b.get_sig_y(),
                    w_avg=b.get_avg_w(), w_sig=b.get_sig_w(), phi_avg=b.get_avg_phi(), phi_sig=b.get_sig_phi(),
                    ex=b.get_emittance_x(), ey=b.get_emittance_y(), ez=b.get_emittance_z())


def compare(a, b, rtol):
    """Per-particle comparison of two beam states; returns a report dict.  Coordinates of particles
    alive in both are compared with a tolerance relative to the rms extent of each coordinate
    (phi: relative to its own magnitude is meaningless for an absolute phase of ~1e3 rad)."""
    rep = {}
    rep["loss_equal"] = bool(np.array_equal(a["loss"], b["loss"]))
    rep["lloss_equal"] = bool(np.array_equal(a["lloss"], b["lloss"]))
    rep["n_loss"] = (int((a["loss"] != 0).sum()), int((b["loss"] != 0).sum()))
    rep["n_lloss"] = (int((a["lloss"] != 0).sum()), int((b["lloss"] != 0).sum()))
    rep["loss_mismatch"] = int((a["loss"] != b["loss"]).sum())
    both = (a["loss"] == 0) & (b["loss"] == 0)
    # "good" in the reference's sense (beam_kernel.cu filters, loss == 0 && lloss == 0): the particles that enter
    # its moments, emittances and the space-charge deposit / kick.  Particles flagged lloss left the RF bucket;
    # they are still tracked, but their trajectories amplify rounding differences without bound.
    good = both & (a["lloss"] == 0) & (b["lloss"] == 0)
    worst = worst_rel = worst_rel_good = p999_rel = median_rel = 0.0
    for k in ("x", "xp", "y", "yp", "phi", "w"):
        va, vb = a[k][both], b[k][both]
        if len(va) == 0:
            continue
        scale = max(np.std(va), 1e-300) if k != "w" else max(np.abs(va).max(), 1e-300)
        if k == "phi":
            scale = max(np.std(va), 1.0)
        d = np.abs(va - vb)
        err = d.max() / scale
        rep["err_" + k] = float(err)
        worst = max(worst, err)
        # the north star's measure: relative to the coordinate's own magnitude (floored at its rms extent,
        # since x, x', y, y' pass through zero)
        relv = d / np.maximum(np.abs(vb), max(np.std(vb), 1e-300))
        rel = relv.max()
        rep["rel_" + k] = float(rel)
        worst_rel = max(worst_rel, rel)
        gsel = good[both]
        if gsel.any():
            worst_rel_good = max(worst_rel_good, float(relv[gsel].max()))
        p999_rel = max(p999_rel, float(np.quantile(relv, 0.999)))
        median_rel = max(median_rel, float(np.median(relv)))
    rep["worst"] = float(worst)
    rep["worst_rel"] = float(worst_rel)
    rep["worst_rel_good"] = float(worst_rel_good)
    rep["p999_rel"] = float(p999_rel)      # 99.9th percentile over the surviving particles, worst coordinate
    rep["median_rel"] = float(median_rel)
    rep["n_good"] = int(good.sum())
    rep["ok"] = rep["loss_equal"] and worst <= rtol
    return rep


# ---- reference CUDA driver (Oracle A) -----------------------------------------------------------
def read_beam_file(path):
    with open(path, "rb") as f:
        assert f.read(4) == b"HPSB"
        ver, n = struct.unpack("<IQ", f.read(12))
        d = {}
        for k in ("x", "xp", "y", "yp", "phi", "w"):
            d[k] = np.frombuffer(f.read(8 * n), dtype=np.float64).copy()
        d["loss"] = np.frombuffer(f.read(4 * n), dtype=np.uint32).copy()
        d["lloss"] = np.frombuffer(f.read(4 * n), dtype=np.uint32).copy()
        rest = f.read()
        if len(rest) >= 4:
            (ns,) = struct.unpack("<I", rest[:4])
            d["stats"] = np.frombuffer(rest[4:4 + 8 * ns], dtype=np.float64).copy()
    return d


def write_beam_file(path, d):
    n = len(d["x"])
    with open(path, "wb") as f:
        f.write(b"HPSB" + struct.pack("<IQ", 1, n))
        for k in ("x", "xp", "y", "yp", "phi", "w"):
            f.write(np.ascontiguousarray(d[k], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(d["loss"], dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(d["lloss"], dtype=np.uint32).tobytes())
        f.write(struct.pack("<I", 0))


REF_STATS = ["loss", "lloss", "good", "x_avg", "x_sig", "y_avg", "y_sig", "phi_avg", "phi_sig", "w_avg", "w_sig",
             "ex", "ey", "ez", "phi_avg_r", "phi_sig_r", "ref_w", "ref_phi", "freq", "current", "design_w"]


def run_reference(workdir, n, first, last, sc=True, double=True, nr=32, nz=128, interval=0.025, adj_bunch=3,
                  cutoff_w=0.8, remesh=0.02, seed=SEED, beam_in=None, design_w=None, freq=FREQ, repeats=1, timeout=900,
                  mesh_cutoff_w=None, extra=()):
    """Runs oracle/_ref/ref_driver{,_dp} (the reference's own CUDA code) and returns (state, log)."""
    exe = REF_DP if double else REF_SP
    os.makedirs(workdir, exist_ok=True)
    out = os.path.join(workdir, "ref_out.bin")
    cfg = [f"lattice {LATTICE}", f"n {n}", f"mass {MASS}", f"charge {CHARGE}", f"current {CURRENT}",
           f"sc {1 if sc else 0}", f"nr {nr}", f"nz {nz}", f"interval {interval}", f"adj_bunch {adj_bunch}",
           f"adj_bunch_cutoff_w {cutoff_w}", f"remesh_threshold {remesh}", f"first {first}", f"last {last}",
           f"beam_out {out}", f"repeats {repeats}"]
    if mesh_cutoff_w is not None:
        cfg.append(f"mesh_cutoff_w {float(mesh_cutoff_w)!r}")
    if beam_in is None:
        cfg += ["gen dc", f"seed {seed}", f"freq {FREQ}", f"ref_w {REF_W}", f"ref_phi {PHI_OFFSET!r}", f"shift_phi {PHI_OFFSET!r}"]
        for k, v in DC.items():
            cfg.append(f"{k} {v!r}")
    else:
        cfg += [f"beam_in {beam_in}", f"freq {float(freq)!r}"]
        if design_w is not None:
            cfg.append(f"design_w {float(design_w)!r}")  # float(): repr(np.float64) is not a C number
    cfg += list(extra)
    path = os.path.join(workdir, "ref.cfg")
    with open(path, "w") as f:
        f.write("\n".join(cfg) + "\n")
    p = subprocess.run([exe, path], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"ref_driver failed ({p.returncode}): {p.stdout[-2000:]} {p.stderr[-2000:]}")
    return read_beam_file(out), p.stdout


def read_monitor_file(path, cap):
    """17 x cap doubles written by ref_driver's `monitor_out` (order of hps_monitor_field)."""
    a = np.fromfile(path, dtype=np.float64).reshape(len(K.MON_FIELDS), cap)
    return {f: a[i].copy() for i, f in enumerate(K.MON_FIELDS)}


def run_reference_cfg(workdir, n, beam_lines, first, last, sc=True, double=True, extra=(), timeout=900):
    """ref_driver with an explicit beam recipe (`beam_lines`: config lines such as "gen waterbag", "seed 1", ...)."""
    exe = REF_DP if double else REF_SP
    os.makedirs(workdir, exist_ok=True)
    out = os.path.join(workdir, "ref_out.bin")
    cfg = [f"lattice {LATTICE}", f"n {n}", f"mass {MASS}", f"charge {CHARGE}", f"current {CURRENT}", f"sc {1 if sc else 0}",
           "nr 32", "nz 128", "interval 0.025", "adj_bunch 3", "adj_bunch_cutoff_w 0.8", "remesh_threshold 0.02",
           f"first {first}", f"last {last}", f"beam_out {out}"] + list(beam_lines) + list(extra)
    path = os.path.join(workdir, "ref.cfg")
    with open(path, "w") as f:
        f.write("\n".join(cfg) + "\n")
    p = subprocess.run([exe, path], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"ref_driver failed ({p.returncode}): {p.stdout[-2000:]} {p.stderr[-2000:]}")
    return read_beam_file(out), p.stdout
