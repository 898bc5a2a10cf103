"""Flat element table ("lattice") for the B200 tracking engine.

`load_from_dbs` restates GenerateBeamLine (reference src/beamline/src/init.cpp:20-429):
for every attached database it walks `<db>.linac` (a UNION view, ordered by
view_index) and pulls the `*_model` columns of each element.  Two details matter
for bit-identical inputs:

* the reference reads every value through `sqlite3_get_table`, i.e. as TEXT
  (REAL columns rendered with "%!.15g"), and converts with `atof`
  (reference src/util/src/sql_utility.cpp:100-112).  We therefore select
  `CAST(col AS TEXT)` and parse with the same semantics (`_atof`), so a REAL
  such as 0.0039883162242155224 becomes 0.00398831622421552 exactly as it does
  in the reference.
* the reference writes `model_index` back into the db while loading
  (init.cpp:68-73); this loader opens the databases read-only and never mutates
  them -- the model index of an element is simply its position in the table.

The table is a numpy structured array with the memory layout of `hps_elem`
(include/hpsim_b200.h) so it can be handed to `hps_lattice_create` unchanged
and saved/loaded as a small binary file (tests/golden/*.lat).
"""
from __future__ import annotations

import os
import re
import sqlite3
import struct
from typing import Iterable, Sequence

import numpy as np

NPARAM = 26
NAME_LEN = 32

ELEM_DTYPE = np.dtype(
    [("type", "<u4"), ("flags", "<u4"), ("name", f"S{NAME_LEN}"), ("p", "<f8", (NPARAM,))]
)
assert ELEM_DTYPE.itemsize == 8 + NAME_LEN + 8 * NPARAM

DRIFT, QUAD, RFGAP_DTL, RFGAP_CCL, ROTATION, CAPERTURE, RAPERTURE, DIAGNOSTICS, BUNCHER, DIPOLE, SPCH_COMP, STEERER = range(1, 13)
FLAG_ON = 1
FLAG_MONITOR = 2

TYPE_NAMES = {
    DRIFT: "Drift", QUAD: "Quad", RFGAP_DTL: "RFGap-DTL", RFGAP_CCL: "RFGap-CCL",
    ROTATION: "Rotation", CAPERTURE: "ApertureC", RAPERTURE: "ApertureR",
    DIAGNOSTICS: "Diagnostics", BUNCHER: "Buncher", DIPOLE: "Dipole",
    SPCH_COMP: "SpchComp", STEERER: "Steerer",
}  # the reference's GetType() strings (beamline_element.cpp constructors)

MAGIC = b"HPSL"
VERSION = 1

_NUM_RE = re.compile(r"^[ \t\n\v\f\r]*([+-]?(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|inf(?:inity)?|nan))", re.I)


def _atof(text) -> float:
    """C `atof`: longest valid numeric prefix, 0.0 when there is none / NULL."""
    if text is None:
        return 0.0
    m = _NUM_RE.match(str(text))
    return float(m.group(1)) if m else 0.0


def _atoi(text) -> int:
    if text is None:
        return 0
    m = re.match(r"^\s*([+-]?\d+)", str(text))
    return int(m.group(1)) if m else 0


class _Db:
    def __init__(self, conn: sqlite3.Connection):
        self.conn = conn

    def one(self, sql: str):
        """GetDataFromDB: first column of first row as text ('' if none)."""
        row = self.conn.execute(sql).fetchone()
        return None if row is None else row[0]


def _txt(col: str) -> str:
    return f"CAST({col} AS TEXT)"


def open_dbs(db_paths: Sequence[str], names: Sequence[str] | None = None) -> tuple[sqlite3.Connection, list[str]]:
    """Open the first db read-only as `main` and attach the rest (DBConnection +
    AttachDB, reference src/util/src/sql_utility.cpp:211-258)."""
    if names is None:
        names = [os.path.splitext(os.path.basename(p))[0] for p in db_paths]
    conn = sqlite3.connect(f"file:{db_paths[0]}?mode=ro", uri=True)
    dbs = ["main"]
    for path, name in zip(db_paths[1:], names[1:]):
        conn.execute(f"attach 'file:{path}?mode=ro' as {name}")
        dbs.append(name)
    return conn, dbs


# number of meaningful parameter slots per element type (for printing)
NPARAM_USED = {DRIFT: 2, QUAD: 3, RFGAP_DTL: 25, RFGAP_CCL: 25, ROTATION: 1, CAPERTURE: 1, RAPERTURE: 4,
               DIAGNOSTICS: 0, BUNCHER: 4, DIPOLE: 9, SPCH_COMP: 1, STEERER: 2}

_TABLE_OF = {"quad": "quad", "drift": "drift", "dtl_gap": "rf_gap", "ccl_gap": "rf_gap", "rotation": "rotation",
             "caperture": "caperture", "raperture": "raperture", "diagnostics": "diagnostics",
             "buncher": "buncher", "dipole": "dipole", "spch_comp": "spch_comp", "steerer": "steerer"}


def _new_elem(etype, name):
    e = np.zeros((), dtype=ELEM_DTYPE)
    e["type"] = etype
    e["name"] = str(name).encode()[: NAME_LEN - 1]
    return e


def _read_element(conn: sqlite3.Connection, q: _Db, db: str, name, view_index, mtype):
    """One row of `linac` -> element record (GenerateBeamLine's per-type blocks, reference src/beamline/src/init.cpp:75-429;
    the same columns PVObserver::UpdateModel re-reads, src/server/src/pv_observer.cpp:92-364).  None for unknown types."""
    new = _new_elem

    def col(table, column, _vi=view_index, _db=db):
        return q.one(f"select {_txt(column)} from {_db}.{table} where view_index = {_vi}")

    if mtype == "quad":
        e = new(QUAD, name)
        e["p"][2] = _atof(col("quad", "gradient_model"))
        e["p"][0] = _atof(col("quad", "length_model"))
        e["p"][1] = _atof(col("quad", "aperture_model"))
        if _atoi(col("quad", "monitor")) > 0:
            e["flags"] |= FLAG_MONITOR
    elif mtype == "drift":
        e = new(DRIFT, name)
        e["p"][0] = _atof(col("drift", "length_model"))
        e["p"][1] = _atof(col("drift", "aperture_model"))
    elif mtype in ("dtl_gap", "ccl_gap"):
        e = new(RFGAP_DTL if mtype == "dtl_gap" else RFGAP_CCL, name)
        key = "cell_id" if mtype == "dtl_gap" else "tank_id"
        ttf = conn.execute(
            "select " + ", ".join(_txt("t1." + c) for c in (
                "beta_g", "beta_min", "ta0", "ta1", "ta2", "ta3", "ta4", "ta5",
                "sa1", "sa2", "sa3", "sa4", "sa5"))
            + f" from {db}.transit_time_factor t1 join {db}.rf_gap t2 on "
            f"t1.module_id = t2.module_id and t1.tank_id = t2.{key} "
            f"where t2.view_index = {view_index}").fetchone()
        p = e["p"]
        p[12:25] = [_atof(v) for v in ttf]
        p[0] = _atof(col("rf_gap", "length_model"))
        p[1] = _atof(col("rf_gap", "frequency_model"))
        p[2] = _atof(q.one(
            f"select {_txt('cell_length_betalambda_design')} from {db}.rf_module where id = "
            f"(select module_id from {db}.rf_gap where view_index = {view_index})"))
        p[3] = _atof(col("rf_gap", "amplitude_model"))
        p[4] = _atof(col("rf_gap", "ref_phase_model"))
        p[5] = _atof(col("rf_gap", "beam_phase_shift_model"))
        p[6] = _atof(col("rf_gap", "energy_out_model"))
        p[7] = _atof(col("rf_gap", "beta_center_model"))
        p[8] = _atof(col("rf_gap", "dg_model"))
        p[9] = _atof(col("rf_gap", "t_model"))
        p[10] = _atof(col("rf_gap", "tp_model"))
        p[11] = _atof(col("rf_gap", "sp_model"))
    elif mtype == "rotation":
        e = new(ROTATION, name)
        e["p"][0] = _atof(col("rotation", "angle_model"))
    elif mtype == "caperture":
        e = new(CAPERTURE, name)
        e["p"][0] = _atof(col("caperture", "aperture_model"))
        if _atoi(col("caperture", "in_out_model")) > 0:
            e["flags"] |= FLAG_ON
    elif mtype == "raperture":
        e = new(RAPERTURE, name)
        e["p"][0] = _atof(col("raperture", "aperture_xl_model"))
        e["p"][1] = _atof(col("raperture", "aperture_xr_model"))
        e["p"][2] = _atof(col("raperture", "aperture_yt_model"))
        e["p"][3] = _atof(col("raperture", "aperture_yb_model"))
        if _atoi(col("raperture", "in_out_model")) > 0:
            e["flags"] |= FLAG_ON
    elif mtype == "diagnostics":
        e = new(DIAGNOSTICS, name)
        if _atoi(col("diagnostics", "monitor")) > 0:
            e["flags"] |= FLAG_MONITOR
    elif mtype == "buncher":
        e = new(BUNCHER, name)
        e["p"][0] = _atof(col("buncher", "voltage_model"))
        e["p"][1] = _atof(col("buncher", "frequency_model"))
        e["p"][2] = _atof(col("buncher", "phase_model"))
        e["p"][3] = _atof(col("buncher", "aperture_model"))
        if _atoi(col("buncher", "on_off")) > 0:
            e["flags"] |= FLAG_ON
    elif mtype == "dipole":
        e = new(DIPOLE, name)
        for k, c in enumerate(("rho_model", "angle_model", "edge_angle1_model",
                               "edge_angle2_model", "half_gap_model", "k1_model",
                               "k2_model", "field_index_model", "kenergy_model")):
            e["p"][k] = _atof(col("dipole", c))
    elif mtype == "spch_comp":
        e = new(SPCH_COMP, name)
        e["p"][0] = _atof(col("spch_comp", "fraction_model"))
    elif mtype == "steerer":
        e = new(STEERER, name)
        e["p"][0] = _atof(col("steerer", "bl_h_model"))
        e["p"][1] = _atof(col("steerer", "bl_v_model"))
    else:
        return None  # GenerateBeamLine ignores unknown model types
    return e


def load_from_conn(conn: sqlite3.Connection, dbs: Iterable[str], write_model_index: bool = False) -> np.ndarray:
    """`write_model_index=True` also records each element's model index in its table row, as
    GenerateBeamLine does (init.cpp:68-73) -- only ever used on a private in-memory copy."""
    q = _Db(conn)
    out = []
    for db in dbs:
        rows = conn.execute(f"select name, {_txt('view_index')}, model_type from {db}.linac").fetchall()
        for name, view_index, mtype in rows:
            e = _read_element(conn, q, db, name, view_index, mtype)
            if e is None:
                continue
            if write_model_index:
                conn.execute(f"update {db}.{_TABLE_OF[mtype]} set model_index = {len(out)} where view_index = {view_index}")
            out.append(e)
    return np.array(out, dtype=ELEM_DTYPE)


def reload_elements(conn: sqlite3.Connection, dbs: Iterable[str], table: np.ndarray, elem_names: Iterable[str]) -> np.ndarray:
    """Re-reads only the named elements (what PVObserver::UpdateModel does for the elements attached to a PV,
    reference src/server/src/pv_observer.cpp:18-23) and returns an updated copy of `table`."""
    q = _Db(conn)
    out = table.copy()
    index = {n.decode(): i for i, n in enumerate(table["name"])}
    for nm in elem_names:
        i = index.get(nm)
        if i is None:
            continue
        for db in dbs:
            row = conn.execute(f"select {_txt('view_index')}, model_type from {db}.linac where name = ?", (nm,)).fetchone()
            if row is None:
                continue
            e = _read_element(conn, q, db, nm, row[0], row[1])
            if e is not None:
                out[i] = e
            break
    return out


def pv_elements(conn: sqlite3.Connection, dbs: Iterable[str], pv: str) -> list[str]:
    """Names of the beamline elements a PV drives: InitPVObserverList (reference src/server/src/init_pv_observer_list.cpp:8-121)
    -- the elements of `channel_list` that name the PV in channel1..4; an rf_module stands for all of its gaps; a master
    phase PV (value_type rf_ph_master) stands for the gaps of every module whose phase channel it steers."""
    names: list[str] = []
    for db in dbs:
        try:
            vt = conn.execute(f"select value_type from {db}.epics_channel where lcs_name = ?", (pv,)).fetchone()
            elems = conn.execute(f"select model_type, name from {db}.channel_list where channel1 = ? or channel2 = ? or channel3 = ? "
                                 f"or channel4 = ?", (pv, pv, pv, pv)).fetchall()
        except sqlite3.Error:
            continue
        if vt is None:
            continue
        for mtype, name in elems:
            if mtype == "rf_module":
                names += [r[0] for r in conn.execute(
                    f"select g.name from {db}.rf_gap g join {db}.rf_module m on m.id = g.module_id where m.name = ?", (name,))]
            else:
                names.append(name)
    return names


def load_from_dbs(db_paths: Sequence[str], names: Sequence[str] | None = None) -> np.ndarray:
    conn, dbs = open_dbs(db_paths, names)
    try:
        return load_from_conn(conn, dbs)
    finally:
        conn.close()


def save(path: str, table: np.ndarray) -> None:
    table = np.ascontiguousarray(table, dtype=ELEM_DTYPE)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", VERSION, len(table)))
        f.write(table.tobytes())


def load(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        head = f.read(12)
        if head[:4] != MAGIC:
            raise ValueError(f"{path}: not a lattice file")
        version, n = struct.unpack("<II", head[4:])
        if version != VERSION:
            raise ValueError(f"{path}: lattice version {version} != {VERSION}")
        data = f.read()
    table = np.frombuffer(data, dtype=ELEM_DTYPE, count=n).copy()
    return table


def names(table: np.ndarray) -> list[str]:
    return [n.decode() for n in table["name"]]


def index_of(table: np.ndarray, name: str) -> int:
    """BeamLine::GetElementModelIndex (reference src/beamline/src/beamline.cpp:70-77)."""
    low = name.lower()
    for i, n in enumerate(names(table)):
        if n.lower() == low:
            return i
    raise KeyError(name)


def count_kicks(table: np.ndarray, first: int, last: int, interval: float) -> int:
    """Number of space-charge kicks simulate(first,last) issues with SC on
    (drift: ceil(L/interval), quad/gap/dipole: 1; simulation_engine.cu:125-373)."""
    import math
    k = 0
    for e in table[first:last + 1]:
        t = int(e["type"])
        if t == DRIFT:
            k += int(math.ceil(e["p"][0] / interval))
        elif t in (QUAD, RFGAP_DTL, RFGAP_CCL, DIPOLE):
            k += 1
    return k


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="dump HPSim sqlite linac dbs to a flat lattice file")
    ap.add_argument("out")
    ap.add_argument("dbs", nargs="+")
    a = ap.parse_args()
    t = load_from_dbs(a.dbs)
    save(a.out, t)
    print(f"{len(t)} elements -> {a.out}")
