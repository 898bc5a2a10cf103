"""HPSim -- host-side mirror of the reference's CPython extension module `HPSim`
(reference src/pywrapper/src/hpsim_module.cpp:398-417): the five types `Beam`, `BeamLine`,
`DBConnection`, `SpaceCharge`, `Simulator` and the module functions `set_gpu`, `set_db_epics`,
`set_db_model`, `get_db_epics`, `get_db_model`, with the same method names, argument meaning,
units (m, rad, MeV, MHz, A) and default options.  Every compute call goes through the C ABI of
include/hpsim_b200.h into the sm_100a library; there is no CPU path in here.

Differences from the reference that a user can observe:
  * errors raise Python exceptions (the reference prints to stderr and returns NULL / exits);
  * the SQLite files are never mutated: DBConnection works on a private in-memory copy, so
    GenerateBeamLine's model_index write-back (src/beamline/src/init.cpp:68-73) and PV writes stay
    local to the process;
  * `Simulator(beam, beamline, spch)` keeps references to its arguments (the reference keeps raw
    pointers, wrap_simulator.cpp:44-51).
"""
from __future__ import annotations

import ctypes as C
import os
import sqlite3

import numpy as np

from . import _capi as K
from . import lattice as L

_backend = None
_device = 0
print_timing = True  # Simulator.simulate prints the reference's "Whole simulation time" line
_ctx_cache = {}


def _use_library(cdll):
    """Test hook: route this module through another implementation of the same C ABI."""
    global _backend
    _backend = cdll  # contexts are cached per (library, device), so nothing has to be dropped here


def _lib():
    return _backend if _backend is not None else K.load()


class _Ctx:
    def __init__(self, lib, device):
        self.lib = lib
        self.h = lib.hps_ctx_create(device)
        if not self.h:
            raise K.HpsError(K.ERR_CUDA, (lib.hps_last_error(None) or b"").decode())

    def check(self, rc):
        if rc != K.OK:
            raise K.HpsError(rc, (self.lib.hps_last_error(self.h) or b"").decode())


def _ctx():
    lib = _lib()
    key = (id(lib), _device)
    if key not in _ctx_cache:
        _ctx_cache[key] = _Ctx(lib, _device)
    return _ctx_cache[key]


def set_gpu(gpu_id: int) -> None:
    """SetGPU (reference src/beamline/src/init.cpp:12-18): objects created afterwards live on this device."""
    global _device
    _device = int(gpu_id)


def comm_unique_id() -> bytes:
    """128-byte NCCL id made on rank 0; the caller distributes it (no reference counterpart)."""
    buf = C.create_string_buffer(128)
    lib = _lib()
    if lib.hps_comm_unique_id(buf) != K.OK:
        raise K.HpsError(K.ERR_COMM, "cannot create an NCCL unique id")
    return buf.raw


def set_math(mode: str) -> None:
    """Arithmetic of the fused segment kernel on the current device: "reference" (the reference's expression
    order with IEEE-exact division / sqrt), "fast" or "reference_nvcc"; see hps_ctx_set_math in
    include/hpsim_b200.h.  No reference counterpart."""
    c = _ctx()
    c.check(c.lib.hps_ctx_set_math(c.h, {"reference": 0, "fast": 1, "reference_nvcc": 2}[mode]))


def get_math() -> str:
    c = _ctx()
    return {0: "reference", 1: "fast", 2: "reference_nvcc"}[c.lib.hps_ctx_get_math(c.h)]


def init_sharding(rank: int, nranks: int, unique_id: bytes) -> str:
    """Particle-shard every Beam created afterwards over `nranks` processes (one GPU each, SPMD:
    all ranks must issue the same HPSim calls in the same order).  Returns the exchange mode."""
    c = _ctx()
    c.check(c.lib.hps_ctx_comm_init(c.h, int(rank), int(nranks), C.create_string_buffer(unique_id, 128)))
    return {0: "single", 1: "p2p", 2: "nccl"}[c.lib.hps_ctx_comm_mode(c.h)]


# ------------------------------------------------------------------------------------------- DB
class DBConnection:
    """reference src/util/inc/sql_utility.h:26-48 + wrap_dbconnection.cpp:21-149"""

    def __init__(self, db_path: str):
        self.dbs = ["main"]
        self.db_addrs = [db_path]
        self.libs = []
        self._keep = []
        self.conn = sqlite3.connect(":memory:", uri=True)
        self.conn.enable_load_extension(True)
        src = sqlite3.connect(f"file:{db_path}?mode=ro", uri=True)
        src.backup(self.conn)
        src.close()

    def load_lib(self, path: str) -> None:
        self.conn.load_extension(path)
        self.libs.append(path)

    def attach_db(self, path: str, name: str) -> None:
        # private shared-cache in-memory copy: Python's backup() can only fill a connection's "main"
        uri = f"file:hpsim_{id(self)}_{name}?mode=memory&cache=shared"
        mem = sqlite3.connect(uri, uri=True)
        src = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
        src.backup(mem)
        src.close()
        self._keep.append(mem)  # the copy lives as long as one connection to it is open
        self.conn.execute(f"attach '{uri}' as {name}")
        self.dbs.append(name)
        self.db_addrs.append(path)

    def print_dbs(self) -> None:
        for n, a in zip(self.dbs, self.db_addrs):
            print(f"{n} -> {a}")

    def print_libs(self) -> None:
        if not self.libs:
            print("No lib is loaded for this DBConnection.")
        for l in self.libs:
            print(l)

    def clear_model_index(self) -> None:
        for db in self.dbs:
            for (tbl,) in self.conn.execute(f"select name from {db}.sqlite_master where type = 'table'").fetchall():
                cols = [r[1] for r in self.conn.execute(f"pragma {db}.table_info({tbl})")]
                if "model_index" in cols:
                    self.conn.execute(f"update {db}.{tbl} set model_index = NULL")

    def get_epics_channels(self):
        out = []
        for db in self.dbs:
            out += [r[0] for r in self.conn.execute(f"select lcs_name from {db}.epics_channel")]
        return out


class BeamLine:
    """reference src/beamline/inc/beamline.h:11-34 + wrap_beamline.cpp:24-116.
    `BeamLine(dbconn)` runs GenerateBeamLine; `BeamLine(table=...)` takes a flat element table."""

    def __init__(self, dbconn: DBConnection | None = None, table: np.ndarray | None = None):
        self._c = _ctx()
        if table is None:
            if not isinstance(dbconn, DBConnection):
                raise TypeError("BeamLine constructor needs a DBConnection!")
            table = L.load_from_conn(dbconn.conn, dbconn.dbs, write_model_index=True)
        self.table = np.ascontiguousarray(table, dtype=L.ELEM_DTYPE)
        self.h = self._c.lib.hps_lattice_create(self._c.h, K.ptr(self.table), len(self.table))
        if not self.h:
            raise K.HpsError(K.ERR_ARG, "lattice creation failed")

    def __del__(self):
        try:
            self._c.lib.hps_lattice_destroy(self.h)
        except Exception:
            pass

    def __len__(self):
        return len(self.table)

    def index(self, name: str) -> int:
        i = self._c.lib.hps_lattice_index(self.h, name.encode())
        if i < 0:
            raise KeyError(f"BeamLineElementNotFound: {name}")  # reference: BeamLineElementNotFoundException
        return int(i)

    def _refresh(self, table: np.ndarray) -> int:
        """push changed element records to the engine (the element setters of pv_observer.cpp:92-364)"""
        changed = 0
        for i in np.nonzero(table != self.table)[0]:
            e = K.Elem.from_buffer_copy(table[i].tobytes())
            self._c.check(self._c.lib.hps_lattice_update_elem(self.h, int(i), C.byref(e)))
            changed += 1
        self.table = table
        return changed

    def print_out(self) -> None:
        self.print_range("", "")

    def print_range(self, start: str, end: str) -> None:
        a = self.index(start) if start else 0
        b = self.index(end) if end else len(self.table) - 1
        for i in range(a, b + 1):
            e = self.table[i]
            print(f"{i}\t{e['name'].decode()}\t{L.TYPE_NAMES[int(e['type'])]}\t" +
                  " ".join(f"{v:.10g}" for v in e["p"][:L.NPARAM_USED[int(e['type'])]]))

    def get_element_names(self, start: str = "", end: str = "", type: str = ""):
        a = self.index(start) if start else 0
        b = self.index(end) if end else len(self.table) - 1
        names = [self.table[i]["name"].decode() for i in range(a, b + 1)
                 if not type or L.TYPE_NAMES[int(self.table[i]["type"])] == type]
        return names or None


# ----------------------------------------------------------------------------------------- Beam
class Beam:
    """reference src/beam/inc/beam.h:140-269 + wrap_beam.cpp:26-50, 1387-1429"""

    def __init__(self, mass: float = 0.0, charge: float = 0.0, current: float = 0.0, num: int = 0, file: str = ""):
        self._c = _ctx()
        lib = self._c.lib
        if file:
            cols = _read_beam_file(file)
            mass, charge, current = cols["mass"], cols["charge"], cols["current"]
            num = len(cols["x"])
        if not num:
            raise ValueError("Beam constructor need either a file name or num of particles as arg!")
        self.h = lib.hps_beam_create(self._c.h, int(num), float(mass), float(charge), float(current))
        if not self.h:
            raise K.HpsError(K.ERR_CUDA, (lib.hps_last_error(self._c.h) or b"").decode())
        off, nl = C.c_uint64(), C.c_uint64()
        lib.hps_beam_local_range(self.h, C.byref(off), C.byref(nl))
        self.offset, self.n_local, self.n = off.value, nl.value, int(num)
        self.stride = int(lib.hps_beam_local_stride(self.h))  # local particle j = global offset + j * stride
        if file:
            lib.hps_beam_set_frequency(self.h, cols["freq"])
            self._upload(cols["x"], cols["xp"], cols["y"], cols["yp"], cols["phi"], cols["w"], cols["loss"], cols["lloss"])
            # Beam::InitBeamFromFile: the first row is the reference particle (beam.cpp:287-296)
            lib.hps_beam_set_ref_w(self.h, float(cols["w"][0]))

    def __del__(self):
        try:
            self._c.lib.hps_beam_destroy(self.h)
        except Exception:
            pass

    # --- distributions
    def _upload(self, x, xp, y, yp, phi, w, loss=None, lloss=None):
        arrs = [K.f64(a) for a in (x, xp, y, yp, phi, w)]
        for a in arrs:
            if len(a) != self.n_local:
                raise ValueError("provided distribution size doesn't match number of particles allocated")
        loss, lloss = K.u32(loss), K.u32(lloss)
        self._c.check(self._c.lib.hps_beam_upload(self.h, *[K.ptr(a) for a in arrs], K.ptr(loss), K.ptr(lloss)))

    def set_distribution(self, x, xp, y, yp, phi, w, loss=None, lloss=None):
        # H19: the reference fills lloss from the `loss` list (wrap_beam.cpp:108-113)
        self._upload(x, xp, y, yp, phi, w, loss, loss if lloss is not None else None)

    def set_waterbag(self, ax, bx, ex, ay, by, ey, az, bz, ez, phi, w, freq, seed: int = 0):
        self._c.check(self._c.lib.hps_beam_set_waterbag(self.h, ax, bx, ex, ay, by, ey, az, bz, ez, phi, w, freq, int(seed)))

    def set_dc(self, ax, bx, ex, ay, by, ey, dphi, phi, w, seed: int = 0):
        self._c.check(self._c.lib.hps_beam_set_dc(self.h, ax, bx, ex, ay, by, ey, dphi, phi, w, int(seed)))

    def save_initial_beam(self): self._c.check(self._c.lib.hps_beam_save(self.h, 0))
    def save_intermediate_beam(self): self._c.check(self._c.lib.hps_beam_save(self.h, 1))
    def restore_initial_beam(self): self._c.check(self._c.lib.hps_beam_restore(self.h, 0))
    def restore_intermediate_beam(self): self._c.check(self._c.lib.hps_beam_restore(self.h, 1))

    def print_to(self, file: str, msg: str = "") -> None:
        """Beam::PrintToFile (beam.cpp:670-707): the five labelled header lines the reference's reader looks for
        (`Info`, `Charge`, `Mass`, `Current`, `Frequency`; header values with the default 6 significant digits of an
        ostream), then one row of 15-significant-digit coordinates + the two loss flags per particle."""
        if self.n_local != self.n:
            raise K.HpsError(K.ERR_STATE, "print_to of a particle-sharded beam would write this rank's shard only; gather the arrays "
                             "(Beam.offset / Beam.stride) and write them from one rank")
        d = self._arrays()

        def num(v):  # the reference's ContainOnlyNumbers (utility.cpp:98-110) does not accept a '+' in a row
            return f"{v:.15g}".replace("e+", "e")
        with open(file, "w") as f:
            f.write(f"Info {msg}\n")
            f.write(f"Charge  {self.get_charge():g}\n")
            f.write(f"Mass {self.get_mass():g}\n")
            f.write(f"Current {self._c.lib.hps_beam_get_current(self.h):g}\n")
            f.write(f"Frequency {self.get_frequency():g}\n")
            cols = [d[k] for k in ("x", "xp", "y", "yp", "phi", "w")]
            loss, lloss = d["loss"], d["lloss"]
            for i in range(self.n_local):
                f.write("  ".join(num(c[i]) for c in cols) + f"  {loss[i]}  {lloss[i]}\n")

    def set_ref_w(self, w): self._c.check(self._c.lib.hps_beam_set_ref_w(self.h, float(w)))
    def set_ref_phi(self, phi): self._c.check(self._c.lib.hps_beam_set_ref_phi(self.h, float(phi)))
    def set_frequency(self, f): self._c.check(self._c.lib.hps_beam_set_frequency(self.h, float(f)))
    def get_mass(self): return self._c.lib.hps_beam_get_mass(self.h)
    def get_charge(self): return self._c.lib.hps_beam_get_charge(self.h)
    def get_frequency(self): return self._c.lib.hps_beam_get_frequency(self.h)
    def get_ref_w(self): return self._c.lib.hps_beam_get_ref_w(self.h)
    def get_ref_phi(self): return self._c.lib.hps_beam_get_ref_phi(self.h)
    def get_size(self): return self.n

    def get_current(self):
        """wrap_beam.cpp BeamGetCurrent: current scaled by the surviving fraction"""
        self._update("loss")
        return self._c.lib.hps_beam_get_current(self.h) * (1.0 - self._stat("num_loss") / float(self.n))

    # --- arrays
    def _arrays(self, which=("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")):
        n = self.n_local
        bufs = {k: (np.empty(n, np.uint32) if k in ("loss", "lloss") else np.empty(n, np.float64)) for k in which}
        order = ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")
        self._c.check(self._c.lib.hps_beam_download(self.h, *[K.ptr(bufs.get(k)) for k in order]))
        return bufs

    def _coord(self, name, option):
        if option not in ("good", "lost", "all"):
            raise ValueError(f'invalid option: {option}, pick from ("good", "lost", "all") !')
        if option == "all":
            return self._arrays((name,))[name]
        d = self._arrays((name, "loss"))
        return d[name][d["loss"] == 0] if option == "good" else d[name][d["loss"] != 0]

    def get_x(self, option="good"): return self._coord("x", option)
    def get_xp(self, option="good"): return self._coord("xp", option)
    def get_y(self, option="good"): return self._coord("y", option)
    def get_yp(self, option="good"): return self._coord("yp", option)
    def get_w(self, option="good"): return self._coord("w", option)

    def get_phi(self, option="good", option2="absolute"):
        if option2 not in ("absolute", "relative"):
            raise ValueError('invalid 2nd option, pick from ("absolute", "relative") !')
        if option2 == "absolute":
            return self._coord("phi", option)
        out = np.empty(self.n_local, np.float64)
        self._c.check(self._c.lib.hps_beam_download_phi_relative(self.h, 0, K.ptr(out)))
        if option == "all":
            return out
        loss = self._arrays(("loss",))["loss"]
        return out[loss == 0] if option == "good" else out[loss != 0]

    def get_losses(self, option="t"):
        if option == "t":
            self._update("loss")
            return self._arrays(("loss",))["loss"]
        if option == "l":
            self._update("lloss")
            return self._arrays(("lloss",))["lloss"]
        raise ValueError('get_losses option must be "t" or "l"')

    def get_loss_num(self, option="t"):
        if option == "t":
            self._update("loss")
            return int(self._stat("num_loss"))
        if option == "l":
            self._update("lloss")
            return int(self._stat("num_lloss"))
        self._update("loss"); self._update("lloss"); self._update("good_count")
        return self.n - int(self._stat("num_good"))

    # --- statistics (update sequences of wrap_beam.cpp:900-1230)
    def _update(self, name): self._c.check(self._c.lib.hps_beam_update(self.h, K.UPD[name]))

    def _stat(self, name):
        v = C.c_double()
        self._c.check(self._c.lib.hps_beam_get(self.h, K.STAT[name], C.byref(v)))
        return v.value

    def _tg_lg(self, option, var, kind):
        self._update("loss")
        if option == "transverse_good":
            self._update(f"{kind}_{var}")
            return self._stat(f"{var}_{kind}")
        if option == "longitudinal_good":
            self._update(f"{kind}_{var}_good")
            return self._stat(f"{var}_{kind}_good")
        raise ValueError('option must be "transverse_good" or "longitudinal_good"')

    def get_avg_x(self, option="transverse_good"): return self._tg_lg(option, "x", "avg")
    def get_avg_y(self, option="transverse_good"): return self._tg_lg(option, "y", "avg")
    def get_avg_w(self, option="transverse_good"): return self._tg_lg(option, "w", "avg")
    def get_sig_x(self, option="transverse_good"): return self._tg_lg(option, "x", "sig")
    def get_sig_y(self, option="transverse_good"): return self._tg_lg(option, "y", "sig")
    def get_sig_w(self, option="transverse_good"): return self._tg_lg(option, "w", "sig")

    def get_avg_phi(self, option="absolute"):
        self._update("loss")
        if option == "absolute":
            self._update("avg_phi"); return self._stat("phi_avg")
        if option == "relative":
            self._update("rel_phi"); self._update("avg_phi_r"); return self._stat("phi_avg_r")
        self._update("good_count"); self._update("avg_phi_good"); return self._stat("phi_avg_good")

    def get_sig_phi(self, option="absolute"):
        self._update("loss")
        if option == "absolute":
            self._update("sig_phi"); return self._stat("phi_sig")
        if option == "relative":
            self._update("rel_phi"); self._update("sig_phi_r"); return self._stat("phi_sig_r")
        self._update("good_count"); self._update("avg_phi_good"); self._update("rel_phi_good")
        self._update("sig_phi_r_good"); return self._stat("phi_sig_good")

    def _emit(self, which):
        self._update("loss"); self._update("emittance")
        return self._stat(which)

    def get_emittance_x(self): return self._emit("x_emit")
    def get_emittance_y(self): return self._emit("y_emit")
    def get_emittance_z(self): return self._emit("z_emit")

    def moment_stats(self):
        """Diagnostics of the fused moment chain (hps_beam_moment_stats; no reference counterpart)."""
        out = (C.c_uint64 * 18)()
        err, m = C.c_double(), C.c_double()
        self._c.check(self._c.lib.hps_beam_moment_stats(self.h, out, C.byref(err), C.byref(m)))
        n = max(1, int(out[8]))
        return {"fused": int(out[0]), "fallback": int(out[1]), "deferred": int(out[2]), "pivot_err_max": err.value, "pivot_margin": m.value,
                "fallback_cause": {"list_overflow": int(out[9]), "marking_band": int(out[10]), "wrap_margin": int(out[11])},
                "sweep_timeline_us": {"launches": int(out[8]), "sweep": out[3] / n / 1e3, "fold": out[4] / n / 1e3, "finalise": out[5] / n / 1e3,
                                      "start_skew": out[6] / n / 1e3, "pilot": out[7] / n / 1e3,
                                      "finalise_phases": [out[12 + k] / n / 1e3 for k in range(6)]}}

    def apply_cut(self, coord, lo, hi):
        m = {"x": K.X, "y": K.Y, "w": K.W, "p": K.PHI}
        self._c.check(self._c.lib.hps_beam_cut(self.h, m[coord.lower()], float(lo), float(hi)))

    def translate(self, axis, val):
        m = {"x": K.X, "xp": K.XP, "y": K.Y, "yp": K.YP, "phi": K.PHI, "w": K.W}
        self._c.check(self._c.lib.hps_beam_shift(self.h, m[axis], float(val)))


def _only_numbers(line: str) -> bool:
    """ContainOnlyNumbers (util/src/utility.cpp:98-110): digits, '.', '-', 'e' / 'E' and blanks"""
    t = line.replace(" ", "")
    return bool(t) and all(ch.isdigit() or ch in ".-eE" for ch in t)


def _first_number(line: str) -> float:
    """GetFirstNumberInString (util/src/utility.cpp:31-85): the first signed decimal number of a line"""
    import re
    m = re.search(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?", line)
    return float(m.group(0)) if m else 0.0


def _read_beam_file(path):
    """Beam::InitBeamFromFile / UpdateBeamFromFile (beam.cpp:179-298): header lines are found by keyword, case-insensitive
    (`charge`, `mass`, `current`, `frequency`; the first number on the line is the value) and each of them is mandatory;
    every line made of numbers only is a particle row with at least 6 columns (+ loss, + lloss)."""
    head, rows = {}, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            low = line.lower()
            for key in ("charge", "mass", "current", "frequency"):
                if key in low and any(ch.isdigit() for ch in line):
                    head[key] = _first_number(line)
            if _only_numbers(line):
                rows.append(line.split())
    for key in ("charge", "mass", "current", "frequency"):
        if key not in head:
            raise ValueError(f"Beam input file error: must define {key} value!")  # the reference prints this and exits
    if not rows:
        raise ValueError("Beam input file error: no particle rows")
    a = np.array([[float(v) for v in r[:6]] for r in rows], dtype=np.float64)
    ncol = len(rows[-1])  # like the reference, the width of the last row decides which flag columns are used
    out = {k: a[:, i].copy() for i, k in enumerate(("x", "xp", "y", "yp", "phi", "w"))}
    out["loss"] = np.array([int(float(r[6])) for r in rows], dtype=np.uint32) if ncol > 6 else None
    out["lloss"] = np.array([int(float(r[7])) for r in rows], dtype=np.uint32) if ncol > 7 else None
    out["mass"], out["charge"], out["current"], out["freq"] = head["mass"], head["charge"], head["current"], head["frequency"]
    return out


# --------------------------------------------------------------------------------- SpaceCharge
class SpaceCharge:
    """reference Scheff (src/simulation/inc/space_charge.h:13-167) + wrap_spacecharge.cpp:31-262"""

    def __init__(self, nr: int = 32, nz: int = 64, interval: float = 0.01, adj_bunch: int = 0, type: str = "scheff"):
        if type != "scheff":
            raise ValueError(f"SpaceCharge type: {type} is not supported.")
        self._c = _ctx()
        self.h = self._c.lib.hps_scheff_create(self._c.h, int(nr), int(nz), int(adj_bunch))
        if not self.h:
            raise K.HpsError(K.ERR_ARG, (self._c.lib.hps_last_error(self._c.h) or b"").decode())
        self._set(K.SC_INTERVAL, interval)

    def __del__(self):
        try:
            self._c.lib.hps_scheff_destroy(self.h)
        except Exception:
            pass

    def _set(self, p, v): self._c.check(self._c.lib.hps_scheff_set(self.h, p, float(v)))

    def _get(self, p):
        v = C.c_double()
        self._c.check(self._c.lib.hps_scheff_get(self.h, p, C.byref(v)))
        return v.value

    def set_mesh_size(self, nr, nz): self._c.check(self._c.lib.hps_scheff_set_mesh_size(self.h, int(nr), int(nz)))
    def get_mesh_size(self): return [self._get(K.SC_NR), self._get(K.SC_NZ)]
    def set_interval(self, v): self._set(K.SC_INTERVAL, v)
    def get_interval(self): return self._get(K.SC_INTERVAL)
    def set_adj_bunch(self, v): self._set(K.SC_ADJ_BUNCH, int(v))
    def get_adj_bunch(self): return int(self._get(K.SC_ADJ_BUNCH))
    def set_adj_bunch_cutoff_w(self, v): self._set(K.SC_ADJ_BUNCH_CUTOFF_W, v)
    def get_adj_bunch_cutoff_w(self): return self._get(K.SC_ADJ_BUNCH_CUTOFF_W)
    def set_mesh_size_cutoff_w(self, v): self._set(K.SC_MESH_CUTOFF_W, v)
    def get_mesh_size_cutoff_w(self): return self._get(K.SC_MESH_CUTOFF_W)
    def set_remesh_threshold(self, v): self._set(K.SC_REMESH_THRESHOLD, v)
    def get_remesh_threshold(self): return self._get(K.SC_REMESH_THRESHOLD)


# ------------------------------------------------------------------------------------ PlotData
class PlotData:
    """reference PlotData (src/util/inc/plot_data.h:43-66): one row of beam moments per monitor element, filled by
    Update2DPlotData (src/simulation/src/simulation_engine.cu:34-82) in the online mode.  The reference's Python module
    does not export it (only the C++ `InitEngine(beam, bl, spch, graph_on, plot_data)` takes one); here
    `Simulator.set_plot_data(PlotData(n))` is that call."""

    def __init__(self, size: int):
        self._c = _ctx()
        self.size = int(size)
        self.h = self._c.lib.hps_monitor_create(self._c.h, self.size)
        if not self.h:
            raise K.HpsError(K.ERR_CUDA, (self._c.lib.hps_last_error(self._c.h) or b"").decode())

    def __del__(self):
        try:
            self._c.lib.hps_monitor_destroy(self.h)
        except Exception:
            pass

    def values(self):
        """{member: numpy array of the rows written so far} (PlotDataMember::GetValue), plus "count"."""
        buf = np.zeros((len(K.MON_FIELDS), self.size))
        n = C.c_uint32()
        self._c.check(self._c.lib.hps_monitor_download(self.h, K.ptr(buf), C.byref(n)))
        k = min(int(n.value), self.size)
        out = {f: buf[i, :k].copy() for i, f in enumerate(K.MON_FIELDS)}
        out["count"] = int(n.value)
        return out


# ----------------------------------------------------------------------------------- Simulator
class Simulator:
    """reference SimulationEngine (src/simulation/src/simulation_engine.cpp:34-95) + wrap_simulator.cpp"""

    def __init__(self, beam: Beam, beamline: BeamLine, spch: SpaceCharge | None = None):
        if not isinstance(beam, Beam) or not isinstance(beamline, BeamLine):
            raise TypeError("Simulator constructor needs at least a beam and beamline pointer as its args!")
        self._c = beam._c
        self.beam, self.beamline, self.spch = beam, beamline, spch
        self.space_charge_on = spch is not None
        self.last_run = None
        self._c.check(self._c.lib.hps_engine_reset(self._c.h, beam.h))  # InitEngine -> Init -> Reset

    def reset(self):
        self._c.check(self._c.lib.hps_engine_reset(self._c.h, self.beam.h))

    def set_plot_data(self, plot_data: "PlotData | None"):
        """InitEngine's (graph_on, plot_data) pair: monitor elements append their rows to `plot_data` from now on."""
        self.plot_data = plot_data
        self._c.check(self._c.lib.hps_ctx_set_monitor(self._c.h, plot_data.h if plot_data is not None else None))

    def set_space_charge(self, on_off: str):
        self.space_charge_on = on_off.lower() == "on"
        if self.space_charge_on and self.spch is None:
            raise ValueError("set_space_charge('on') needs a SpaceCharge object")

    def simulate(self, start: str = "", end: str = ""):
        first = self.beamline.index(start) if start else 0
        last = self.beamline.index(end) if end else len(self.beamline) - 1
        rs = K.RunStats()
        sc = self.spch.h if (self.space_charge_on and self.spch is not None) else None
        self._c.check(self._c.lib.hps_simulate(self._c.h, self.beam.h, self.beamline.h, sc, first, last, C.byref(rs)))
        self.last_run = rs.as_dict()
        if print_timing:
            print(f"Whole simulation time: {rs.device_ms / 1e3:g} [sec]")  # util/src/timer.cpp:10-19
        return None


# ------------------------------------------------------------------------ module-level functions
def get_db_epics(pv: str, dbconn: DBConnection) -> str:
    """hpsim_module.cpp GetDbEPICS: value column of epics_channel as text"""
    for db in dbconn.dbs:
        row = dbconn.conn.execute(f"select CAST(value AS TEXT) from {db}.epics_channel where lcs_name = ?", (pv,)).fetchone()
        if row is not None and row[0] is not None:
            return row[0]
    return ""


def get_db_model(record: str, field: str, dbconn: DBConnection) -> str:
    """hpsim_module.cpp GetDbModel: first table holding both the column and the named record"""
    for db in dbconn.dbs:
        for (tbl,) in dbconn.conn.execute(f"select name from {db}.sqlite_master where type = 'table'").fetchall():
            cols = [r[1] for r in dbconn.conn.execute(f"pragma {db}.table_info({tbl})")]
            if field in cols and "name" in cols:
                row = dbconn.conn.execute(f"select CAST({field} AS TEXT) from {db}.{tbl} where name = ?", (record,)).fetchone()
                if row is not None and row[0] is not None:
                    return row[0]
    return ""


def _reload(dbconn: DBConnection, beamline: BeamLine, names=None) -> int:
    """Push what a database write changed down to the engine.  `names`: only these elements are re-read (the reference's
    PVObserver re-reads the elements attached to the PV, pv_observer.cpp:18-23); None re-reads the whole line (~1 s for
    the 6 540-element LANSCE lattice -- tens of thousands of single-column SELECTs)."""
    if names is None:
        return beamline._refresh(L.load_from_conn(dbconn.conn, dbconn.dbs, write_model_index=False))
    return beamline._refresh(L.reload_elements(dbconn.conn, dbconn.dbs, beamline.table, names))


def set_db_epics(pv: str, value: str, dbconn: DBConnection, beamline: BeamLine) -> None:
    """hpsim_module.cpp:56-83 + PVObserver::Update (src/server/src/pv_observer.cpp:18-23): write the
    PV, let the database triggers recompute the *_model columns, push the changed elements down."""
    print(f"set_db_epics(): pv = {pv}, value = {value}")
    for db in dbconn.dbs:
        cur = dbconn.conn.execute(f"update {db}.epics_channel set value = ? where lcs_name = ?", (str(value), pv))
        if cur.rowcount:
            break
    names = L.pv_elements(dbconn.conn, dbconn.dbs, pv)
    _reload(dbconn, beamline, names if names else None)  # a PV nobody listens to in channel_list: re-read everything


def set_db_model(record: str, field: str, value: str, dbconn: DBConnection, beamline: BeamLine) -> None:
    """hpsim_module.cpp:85-267: only the buncher / rf_module / spch_comp / raperture / caperture tables"""
    for db in dbconn.dbs:
        for tbl in ("buncher", "rf_module", "spch_comp", "raperture", "caperture"):
            try:
                cols = [r[1] for r in dbconn.conn.execute(f"pragma {db}.table_info({tbl})")]
            except sqlite3.Error:
                continue
            if field not in cols:
                continue
            if dbconn.conn.execute(f"select id from {db}.{tbl} where name = ?", (record,)).fetchone() is None:
                continue
            dbconn.conn.execute(f"update {db}.{tbl} set {field} = {value} where name = ?", (record,))
            if tbl == "rf_module":  # a module stands for all of its gaps
                names = [r[0] for r in dbconn.conn.execute(
                    f"select g.name from {db}.rf_gap g join {db}.rf_module m on m.id = g.module_id where m.name = ?", (record,))]
            else:
                names = [record]
            _reload(dbconn, beamline, names)
            return
