"""BASELINE configs[0] ("C1", the reference script's own workload): 64 Ki particles, TBDB02 -> 48DT (6 476 elements,
8 894 space-charge kicks), SCHEFF 32x128 -- and its LEBT-only sub-case.  Prints steps/s of the product; run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H
from hpsim_b200 import HPSim
HPSim.print_timing = False
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for first, last in (("TBDB02", "spch_comp_7"), ("TBDB02", "48DT")):
    c = H.Case(H.cuda_lib(), n=n, sc=True)
    c.run(first, last)  # warm-up (plan upload, lazy allocations)
    best = None
    for _ in range(3):
        c.beam.restore_initial_beam()
        t0 = time.time()
        r = c.run(first, last)
        c.beam._c.check(c.lib.hps_ctx_synchronize(c.beam._c.h))
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    nel = c.beamline.index(last) - c.beamline.index(first) + 1
    print(f"C1 n={n} {first}->{last}: {nel} elements, {r['kicks']} kicks, {r['launches']} launches, wall {best*1e3:.1f} ms, "
          f"device {r['device_ms']:.1f} ms, {n*nel/best:.3e} steps/s, loss {c.beam.get_loss_num('t')}")
