"""Diagnostic (GPU box): dump reference-CUDA end states at several end elements / settings."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H

out_dir = os.path.join(H.ROOT, "gpurun_out", "bisect")
os.makedirs(out_dir, exist_ok=True)
lat = H.lattice()
names = [x.decode() for x in lat["name"]]
cases = []
for idx in (130, 131, 132, 133, 134, 136, 140, 150, 170, 200, 260, 330, 400):
    cases.append((f"dtl_to_{idx}", dict(n=4096, first="TBDB02", last=names[idx], sc=True)))
cases.append(("lebt_adj0", dict(n=4096, first="TBDB02", last="spch_comp_7", sc=True, adj_bunch=0)))
cases.append(("lebt_adj0_cut", dict(n=4096, first="TBDB02", last="spch_comp_7", sc=True, adj_bunch=3, cutoff_w=0.5)))
cases.append(("dtl_rep_a", dict(n=4096, first="TBDB02", last="04QM30", sc=True)))
cases.append(("dtl_rep_b", dict(n=4096, first="TBDB02", last="04QM30", sc=True)))
for name, kw in cases:
    st, log = H.run_reference(os.path.join(H.ROOT, "gpurun_out", "bisect_work"), double=True, **kw)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), stats=st["stats"], **{k: st[k] for k in ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")})
    print(name, int((st["loss"] != 0).sum()), int((st["lloss"] != 0).sum()))
