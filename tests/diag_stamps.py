"""diagnostic: per-block phase stamps of the fused space-charge kernel at the last re-mesh kick of a C2 pass"""
import ctypes as C, sys
import numpy as np
import harness as H
from hpsim_b200 import HPSim, _capi as K
HPSim.print_timing = False
lib = K.load()
lib.hps_debug_scheff_stamps.restype = C.c_int
lib.hps_debug_scheff_stamps.argtypes = [C.c_void_p, C.c_void_p]
c = H.Case(lib, n=1 << 20, sc=True)
c.run("TBDB02", "spch_comp_7")
c.run("01QM00U", "02QM10")
buf = np.zeros((1024, 16), dtype=np.uint64)
lib.hps_debug_scheff_stamps(c.beam._c.h, buf.ctypes.data_as(C.c_void_p))
nb = int((buf[:, 0] != 0).sum())
b = buf[:nb].astype(np.int64)
t0 = b[:, 0].min()
names = ["start", "moments", "barrier1", "deposit", "green", "barrier2", "field"]
print("blocks", nb)
for k, nm in enumerate(names):
    col = b[:, k] - t0
    ok = b[:, k] != 0
    print(f"{nm:9s} min {col[ok].min()/1e3:7.2f} med {np.median(col[ok])/1e3:7.2f} max {col[ok].max()/1e3:7.2f} us   (n={int(ok.sum())})")
d = b[:, 3] - b[:, 2]
print("deposit duration per block: min %.2f med %.2f max %.2f us" % (d.min()/1e3, np.median(d)/1e3, d.max()/1e3))
g = (b[:, 4] - b[:, 3])
print("green duration per block: min %.2f med %.2f max %.2f us" % (g.min()/1e3, np.median(g)/1e3, g.max()/1e3))
f = (b[:, 6] - b[:, 5])[b[:, 6] != 0]
print("field duration per block: min %.2f med %.2f max %.2f us" % (f.min()/1e3, np.median(f)/1e3, f.max()/1e3))
m = (b[:, 1] - b[:, 0])
print("moments duration per block: min %.2f med %.2f max %.2f us; slowest block %d" % (m.min()/1e3, np.median(m)/1e3, m.max()/1e3, int(m.argmax())))

for k, nm in enumerate(["gathered", "convolved", "reduced+stored", "ticket"]):
    col = (b[:, 8 + k] - b[:, 5])[b[:, 8 + k] != 0]
    print(f"field sub-phase {nm:15s} (since barrier2) min {col.min()/1e3:6.2f} med {np.median(col)/1e3:6.2f} max {col.max()/1e3:6.2f} us")
