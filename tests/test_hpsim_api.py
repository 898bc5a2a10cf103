"""Host-side mirror of the reference's HPSim module, exercised on the CPU oracle backend
(the same Python code drives the CUDA library; only the .so differs)."""
import numpy as np
import pytest

import harness as H
from hpsim_b200 import HPSim, _capi as K, lattice as L

HPSim.print_timing = False


@pytest.fixture()
def case():
    return H.Case(H.oracle_lib(), n=2048, sc=True)


def test_getters_and_options(case):
    b = case.beam
    assert b.get_size() == 2048 and b.get_mass() == H.MASS and b.get_charge() == H.CHARGE
    assert b.get_frequency() == H.FREQ and b.get_ref_w() == H.REF_W and b.get_ref_phi() == H.PHI_OFFSET
    assert len(b.get_x()) == 2048 and len(b.get_x("lost")) == 0 and len(b.get_x("all")) == 2048  # default is "good" (H19)
    with pytest.raises(ValueError):
        b.get_x("bogus")
    assert b.get_losses().dtype == np.uint32 and b.get_loss_num() == 0 and b.get_loss_num("all") == 0
    rel = b.get_phi("all", "relative")
    assert np.all(np.abs(rel) <= np.pi + 1e-12) and rel[0] == 0.0  # relative to the reference particle


def test_save_restore_and_translate(case):
    b = case.beam
    x0 = b.get_x("all").copy()
    b.translate("x", 0.001)
    x1 = b.get_x("all")
    assert x1[0] == x0[0] and np.allclose(x1[1:], x0[1:] + 0.001)  # particle 0 is never shifted (H1)
    b.restore_initial_beam()
    assert np.array_equal(b.get_x("all"), x0)
    fresh = HPSim.Beam.__new__(HPSim.Beam)
    HPSim._use_library(H.oracle_lib())
    try:
        nb = HPSim.Beam(mass=1.0, charge=1.0, current=0.0, num=8)
        with pytest.raises(K.HpsError):
            nb.restore_initial_beam()  # nothing saved
    finally:
        HPSim._use_library(None)


def test_apply_cut_marks_99999(case):
    b = case.beam
    b.apply_cut("x", -1.0, 1.0)  # everything inside (lo, hi) is cut, except the reference particle
    loss = b.get_losses()
    assert loss[0] == 0 and np.all(loss[1:] == 99999)
    assert b.get_loss_num() == 2047


def test_set_distribution_quirks(case):
    n = 2048
    z = [0.0] * n
    x = [1e-3] * n
    loss = [0] * n
    loss[5] = 7
    case.beam.set_distribution(x, z, z, z, z, [0.75] * n, loss, [0] * n)
    d = case.beam._arrays()
    assert d["x"][0] == 0.0 and d["x"][1] == 1e-3      # reference particle forced on axis
    assert d["loss"][5] == 7 and d["lloss"][5] == 7    # H19: lloss is read from the loss list


def test_simulate_resume_equals_single_run():
    a = H.Case(H.oracle_lib(), n=1024, sc=True)
    a.run("TBDB02", "spch_comp_7")
    b = H.Case(H.oracle_lib(), n=1024, sc=True)
    b.run("TBDB02", "TBQL05V2")
    nxt = b.beamline.get_element_names()[b.beamline.index("TBQL05V2") + 1]
    b.run(nxt, "spch_comp_7")
    sa, sb = a.state(), b.state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k


def test_unknown_element_raises(case):
    with pytest.raises(KeyError):
        case.sim.simulate("TBDB02", "no_such_element")


def test_space_charge_setters():
    HPSim._use_library(H.oracle_lib())
    try:
        s = HPSim.SpaceCharge(nr=16, nz=64)
        assert s.get_interval() == 0.01 and s.get_adj_bunch() == 0 and s.get_remesh_threshold() == 0.05
        assert s.get_mesh_size() == [16.0, 64.0]
        s.set_mesh_size(32, 128); s.set_interval(0.025); s.set_adj_bunch(3)
        s.set_adj_bunch_cutoff_w(0.8); s.set_remesh_threshold(0.02); s.set_mesh_size_cutoff_w(0.0)
        assert s.get_mesh_size() == [32.0, 128.0] and s.get_interval() == 0.025 and s.get_adj_bunch() == 3
        assert s.get_adj_bunch_cutoff_w() == 0.8 and s.get_remesh_threshold() == 0.02
        with pytest.raises(ValueError):
            HPSim.SpaceCharge(nr=8, nz=8, type="picnic")
    finally:
        HPSim._use_library(None)


def test_lattice_counts():
    lat = H.lattice()
    assert len(lat) == 6540
    first, last = L.index_of(lat, "TBDB02"), L.index_of(lat, "48DT")
    assert (first, last) == (62, 6537)
    assert L.count_kicks(lat, first, last, 0.025) == 8894
    assert L.count_kicks(lat, first, L.index_of(lat, "spch_comp_7"), 0.025) == 236
    assert L.count_kicks(lat, L.index_of(lat, "01QM00U"), L.index_of(lat, "04QM30"), 0.025) == 406


def test_oracle_kick_count_matches_planner_formula():
    c = H.Case(H.oracle_lib(), n=256, sc=True)
    run = c.run("TBDB02", "spch_comp_7")
    assert run["kicks"] == 236 and run["elements"] == 68


def test_lattice_loader_against_reference_dbs():
    import os
    db = H.DB_DIR
    t = L.load_from_dbs([os.path.join(db, f) for f in ("tbtd.db", "dtl.db", "trst.db", "ccl.db")])
    g = H.lattice()
    assert len(t) == len(g) and np.array_equal(t, g)


def test_dbconnection_and_beamline_from_db():
    import os
    db = H.DB_DIR
    HPSim._use_library(H.oracle_lib())
    try:
        conn = HPSim.DBConnection(os.path.join(db, "tbtd.db"))
        conn.attach_db(os.path.join(db, "dtl.db"), "dtl")
        conn.clear_model_index()
        bl = HPSim.BeamLine(conn)
        assert bl.get_element_names("TBDB02", "TBDB02") == ["TBDB02"]
        assert len(bl.get_element_names(type="Buncher")) >= 2
        assert HPSim.get_db_model("TDDB01", "frequency_model", conn) != ""
        assert len(conn.get_epics_channels()) > 0
    finally:
        HPSim._use_library(None)


def test_set_db_epics_and_set_db_model_reach_the_lattice():
    """hps.set_db_epics / set_db_model: a PV write fires the database triggers (they need the
    vendored sqlite math extension), the *_model columns are re-read and the changed element is pushed down through
    hps_lattice_update_elem.  PV TDDB001E04 is the amplitude set point of the main buncher TDDB01 (the scan script's PV)."""
    import os
    db = H.DB_DIR
    HPSim._use_library(H.oracle_lib())
    try:
        conn = HPSim.DBConnection(os.path.join(db, "tbtd.db"))
        conn.load_lib(H.SQLITE_EXT)
        conn.clear_model_index()
        bl = HPSim.BeamLine(conn)
        i = bl.index("TDDB01")
        lib = bl._c.lib
        e0 = K.Elem()
        assert lib.hps_lattice_get_elem(bl.h, i, e0) == 0
        v0 = float(HPSim.get_db_epics("TDDB001E04", conn))
        HPSim.set_db_epics("TDDB001E04", str(v0 * 1.05), conn, bl)
        assert abs(float(HPSim.get_db_epics("TDDB001E04", conn)) - v0 * 1.05) < 1e-9
        e1 = K.Elem()
        assert lib.hps_lattice_get_elem(bl.h, i, e1) == 0
        assert e1.p[0] != e0.p[0] and abs(e1.p[0] / e0.p[0] - 1.0) < 0.2     # buncher voltage_model follows the PV
        assert [e1.p[k] for k in range(1, 6)] == [e0.p[k] for k in range(1, 6)]
        HPSim.set_db_epics("TDDB001E04", str(v0), conn, bl)
        e2 = K.Elem()
        lib.hps_lattice_get_elem(bl.h, i, e2)
        assert e2.p[0] == e0.p[0]
        # set_db_model writes a *_model column directly
        f0 = HPSim.get_db_model("TDDB01", "phase_offset_cal", conn)
        HPSim.set_db_model("TDDB01", "phase_offset_cal", "12.5", conn, bl)
        assert float(HPSim.get_db_model("TDDB01", "phase_offset_cal", conn)) == 12.5
        if f0 != "":
            HPSim.set_db_model("TDDB01", "phase_offset_cal", f0, conn, bl)
    finally:
        HPSim._use_library(None)


def test_monitor_rows_on_the_oracle():
    """PlotData / Update2DPlotData through the mirror (HPSim.PlotData, Simulator.set_plot_data) on the CPU oracle: one row per
    monitor quadrupole, restarted by an engine reset; the GPU counterpart is tests/test_gpu_full_linac.py."""
    lib = H.oracle_lib()
    c = H.Case(lib, n=2048, sc=True)
    HPSim._use_library(lib)
    try:
        pd = HPSim.PlotData(16)
    finally:
        HPSim._use_library(None)
    c.sim.set_plot_data(pd)
    c.run("TBDB02", "01QM10")
    v = pd.values()
    assert v["count"] == 11 and len(v["xsig"]) == 11
    assert np.all(v["xsig"] > 0) and np.all(v["zemit"] > 0) and np.all(np.diff(v["loss_ratio"]) >= 0)
    assert np.all(v["loss_local"] >= -6.0)
    c.beam.restore_initial_beam()
    c.run("TBDB02", "01QM03")            # restart upstream: Reset() zeroes monitor_num
    assert pd.values()["count"] == 4
    c.sim.set_plot_data(None)


def test_set_db_epics_rereads_only_the_elements_of_the_pv():
    """set_db_epics re-reads the elements attached to the PV (InitPVObserverList's channel_list relations, reference
    src/server/src/init_pv_observer_list.cpp:8-121) instead of the whole line; the result must equal a full reload."""
    import random
    HPSim._use_library(H.oracle_lib())
    try:
        conn = HPSim.DBConnection(H.DB_FILES[0])
        conn.load_lib(H.SQLITE_EXT)
        for f, nm in zip(H.DB_FILES[1:], ("dtl", "trst", "ccl")):
            conn.attach_db(f, nm)
        conn.clear_model_index()
        bl = HPSim.BeamLine(conn)
        assert L.pv_elements(conn.conn, conn.dbs, "TDDB001E04") == ["TDDB01"]
        random.seed(7)
        for pv in random.sample(conn.get_epics_channels(), 12):
            v = HPSim.get_db_epics(pv, conn)
            try:
                nv = repr(float(v) * 1.013 + 0.01)
            except ValueError:
                continue
            before = bl.table.copy()
            HPSim.set_db_epics(pv, nv, conn, bl)
            full = L.load_from_conn(conn.conn, conn.dbs)
            assert np.array_equal(full, bl.table), pv
            names = L.pv_elements(conn.conn, conn.dbs, pv)
            changed = {bl.table["name"][i].decode() for i in np.nonzero(before != bl.table)[0]}
            assert changed <= set(names), (pv, changed, names[:4])
    finally:
        HPSim._use_library(None)
