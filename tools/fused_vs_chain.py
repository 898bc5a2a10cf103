#!/usr/bin/env python
"""Diagnostic (1 GPU): the fused moment epilogue of k_segment against the stand-alone chain (HPSIM_FUSED_MOMENTS=0) on
the same beam and range -- loss / lloss flags must be identical and the coordinates equal to rounding.  When they are
not, bisects the end element for the first stretch on which the two differ and prints the kick statistics there.

  python tools/fused_vs_chain.py [n_particles] [first] [last]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import harness as H  # noqa: E402
from hpsim_b200 import HPSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
first = sys.argv[2] if len(sys.argv) > 2 else "TBDB02"
last = sys.argv[3] if len(sys.argv) > 3 else "01QM10"
HPSim.print_timing = False


def make(fused):
    os.environ["HPSIM_FUSED_MOMENTS"] = "1" if fused else "0"
    HPSim._ctx_cache.clear()  # the switch is read when the context is created
    return H.Case(H.cuda_lib(), n=n, sc=True)


def run_to(case, end):
    case.beam.restore_initial_beam()
    case.sim.reset() if hasattr(case.sim, "reset") else None
    case.run(first, end)
    return case.state()


def diff(a, b):
    return int((a["lloss"] != b["lloss"]).sum()), int((a["loss"] != b["loss"]).sum())


cf = make(True)
sf = run_to(cf, last)
mf = cf.beam.moment_stats()
keep_f = dict(HPSim._ctx_cache)
cc = make(False)
sc = run_to(cc, last)
dl, dt = diff(sf, sc)
rep = H.compare(sf, sc, 1e-9)
print("N", n, "range", first, last, "lloss diff", dl, "loss diff", dt, "n_lloss", rep["n_lloss"], "worst", rep["worst"], "fused stats",
      {k: mf[k] for k in ("fused", "fallback", "deferred", "pivot_err_max", "fallback_cause") if k in mf})
if dl == 0 and dt == 0:
    sys.exit(0)

# bisect the end element: both runs restart from the saved initial beam, so every probe is independent
names = cf.beamline.get_element_names(first, last)
lo, hi = 0, len(names) - 1  # invariant: equal up to names[lo-1] (or nothing run), different at names[hi]
ctx_c = dict(HPSim._ctx_cache)
while lo < hi:
    mid = (lo + hi) // 2
    HPSim._ctx_cache.clear(); HPSim._ctx_cache.update(keep_f)
    a = run_to(cf, names[mid])
    HPSim._ctx_cache.clear(); HPSim._ctx_cache.update(ctx_c)
    b = run_to(cc, names[mid])
    d = diff(a, b)
    print("  probe", mid, names[mid], d)
    if d[0] or d[1]:
        hi = mid
    else:
        lo = mid + 1
print("first differing end element:", lo, names[lo])
HPSim._ctx_cache.clear(); HPSim._ctx_cache.update(keep_f)
a = run_to(cf, names[lo])
HPSim._ctx_cache.clear(); HPSim._ctx_cache.update(ctx_c)
b = run_to(cc, names[lo])
idx = np.nonzero(a["lloss"] != b["lloss"])[0]
print("differing particles:", len(idx), idx[:16])
for k in ("phi", "w", "x", "y"):
    print(" ", k, "fused", a[k][idx[:6]], "chain", b[k][idx[:6]])
print("  lloss fused", a["lloss"][idx[:6]], "chain", b["lloss"][idx[:6]], "loss", a["loss"][idx[:6]], b["loss"][idx[:6]])
good_f = (a["loss"] == 0) & (a["lloss"] == 0)
good_c = (b["loss"] == 0) & (b["lloss"] == 0)
print("  <phi>_good fused", a["phi"][good_f].mean(), "chain", b["phi"][good_c].mean(), "thr-ish", b["phi"][good_c].mean() + 3 * np.pi)
