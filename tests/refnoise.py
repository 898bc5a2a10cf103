"""Diagnostic (GPU box): reference-vs-reference run-to-run spread and product-vs-reference deviation."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H
from hpsim_b200 import HPSim
HPSim.print_timing = False

def dev(a, b):
    good = (a["loss"] == 0) & (b["loss"] == 0) & (a["lloss"] == 0) & (b["lloss"] == 0)
    out = {"loss_equal": bool(np.array_equal(a["loss"], b["loss"])), "n_good": int(good.sum())}
    for k in ("x", "xp", "y", "yp", "phi", "w"):
        e = np.abs(a[k][good] - b[k][good]) / max(np.std(b[k][good]), 1e-300)
        out[k] = [float(np.median(e)), float(e.max())]
    return out

res = {}
for n, first, last in ((65536, "TBDB02", "spch_comp_7"), (65536, "TBDB02", "04QM30"), (65536, "TBDB02", "48DT")):
    key = f"{n}_{last}"
    w = os.path.join(H.ROOT, "gpurun_out", "noise_work")
    r1, _ = H.run_reference(w, n, first, last, sc=True, double=True)
    r2, _ = H.run_reference(w, n, first, last, sc=True, double=True)
    g = H.Case(H.cuda_lib(), n=n, sc=True)
    g.run(first, last)
    g.beam.get_loss_num("l")
    gs = g.state()
    g2 = H.Case(H.cuda_lib(), n=n, sc=True)
    g2.run(first, last)
    g2.beam.get_loss_num("l")
    res[key] = {"ref_vs_ref": dev(r1, r2), "cuda_vs_ref": dev(gs, r1), "cuda_vs_cuda": dev(gs, g2.state())}
    print(key, json.dumps(res[key]))
json.dump(res, open(os.path.join(H.ROOT, "gpurun_out", "refnoise.json"), "w"), indent=1)
