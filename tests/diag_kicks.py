"""Diagnostic (GPU box): the mid-DTL restart scenario of test_live_reference_single_kicks_in_dtl,
reference vs product vs CPU oracle, with the differing particles listed."""
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
beam_in = os.path.join(tmp, "mid.bin")
H.write_beam_file(beam_in, mid)
names = [x.decode() for x in H.lattice()["name"]]
i0 = names.index("01RG20")
for last in sys.argv[1:] or ["01RG20", "01QM21", "01QM22"]:
    ref, log = H.run_reference(tmp, n, "01RG20", last, sc=True, double=True, beam_in=beam_in, design_w=stats["design_w"], freq=stats["freq"])
    out = {}
    for tag, lib in (("cuda", H.cuda_lib()), ("oracle", H.oracle_lib())):
        c = H.Case(lib, n=n, sc=True, make_beam=False)
        c.beam.set_frequency(stats["freq"])
        c.beam._upload(mid["x"], mid["xp"], mid["y"], mid["yp"], mid["phi"], mid["w"], mid["loss"], mid["lloss"])
        T._set_design_w(c, stats["design_w"])
        r = c.run("01RG20", last)
        c.beam.get_loss_num("l")
        st = c.state()
        rep = H.compare(st, ref, 1e-9)
        dl = np.nonzero(st["loss"] != ref["loss"])[0]
        dll = np.nonzero(st["lloss"] != ref["lloss"])[0]
        rep["run"] = r
        rep["loss_diff_idx"] = dl[:12].tolist()
        rep["loss_diff_codes"] = [(int(st["loss"][i]), int(ref["loss"][i])) for i in dl[:12]]
        rep["lloss_diff_idx"] = dll[:12].tolist()
        rep["lloss_diff"] = [(int(st["lloss"][i]), int(ref["lloss"][i]), float(st["phi"][i]), float(ref["phi"][i]), int(mid["lloss"][i]), int(mid["loss"][i])) for i in dll[:12]]
        out[tag] = rep
    print("DIAG", last, names[i0:names.index(last) + 1], json.dumps(out, default=str))
    print("REFSTATS", json.dumps(dict(zip(H.REF_STATS, ref["stats"].tolist()))))
