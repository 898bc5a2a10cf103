"""Diagnostic (GPU box): restart at 01RG20 from a reference mid-DTL state; which particles differ?"""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H
from hpsim_b200 import HPSim
import test_gpu_parity as T
HPSim.print_timing = False
tmp = tempfile.mkdtemp(prefix="hpsim_diag_")
n = 32768
mid, _ = H.run_reference(tmp, n, "TBDB02", "01QM19", sc=True, double=True)
stats = dict(zip(H.REF_STATS, mid["stats"]))
print("MIDSTATS", json.dumps(stats))
beam_in = os.path.join(tmp, "mid.bin")
H.write_beam_file(beam_in, mid)
for sc in (False, True):
    for last in ("01RG20",):
        ref, log = H.run_reference(tmp, n, "01RG20", last, sc=sc, double=True, beam_in=beam_in, design_w=stats["design_w"], freq=stats["freq"])
        c = H.Case(H.oracle_lib(), n=n, sc=sc, make_beam=False)
        c.beam.set_frequency(stats["freq"])
        c.beam._upload(mid["x"], mid["xp"], mid["y"], mid["yp"], mid["phi"], mid["w"], mid["loss"], mid["lloss"])
        T._set_design_w(c, stats["design_w"])
        c.run("01RG20", last)
        c.beam.get_loss_num("l")
        st = c.state()
        both = (st["loss"] == 0) & (ref["loss"] == 0)
        dphi = np.abs(st["phi"] - ref["phi"]); dw = np.abs(st["w"] - ref["w"]); dx = np.abs(st["xp"] - ref["xp"])
        bad = both & ((dphi > 1e-9) | (dw > 1e-12))
        print("SC", sc, "alive", int(both.sum()), "bad", int(bad.sum()), "max dphi", float(dphi[both].max()), "max dw", float(dw[both].max()), "max dxp", float(dx[both].max()))
        idx = np.nonzero(bad)[0][:8]
        good_in = (mid["loss"] == 0) & (mid["lloss"] == 0)
        pavg = mid["phi"][good_in].mean()
        for i in idx:
            print("  i", int(i), "mid phi-avg", float(mid["phi"][i] - pavg), "mid w", float(mid["w"][i]), "lloss_mid", int(mid["lloss"][i]),
                  "mine phi,w", float(st["phi"][i]), float(st["w"][i]), "ref phi,w", float(ref["phi"][i]), float(ref["w"][i]),
                  "mid x,y", float(mid["x"][i]), float(mid["y"][i]))
        print("  bad among good_in", int((bad & good_in).sum()), "bad among lloss_in", int((bad & (mid["lloss"] != 0)).sum()), "n lloss_in", int((mid["lloss"] != 0).sum()))
        print("  w[0] mid", float(mid["w"][0]), "phi[0] mid", float(mid["phi"][0]), "mine w0", float(st["w"][0]), "ref w0", float(ref["w"][0]))
