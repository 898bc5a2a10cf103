"""CLI parity report (used under gpurun): CUDA product vs CPU oracle vs the reference's own CUDA."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def main():
    from hpsim_b200 import HPSim
    HPSim.print_timing = False
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    first = sys.argv[2] if len(sys.argv) > 2 else "TBDB02"
    last = sys.argv[3] if len(sys.argv) > 3 else "spch_comp_7"
    sc = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
    use_ref = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
    out = {}
    t = time.time()
    g = H.Case(H.cuda_lib(), n=n, sc=sc)
    out["cuda_run"] = g.run(first, last)
    gs = g.state()
    out["cuda_wall_s"] = time.time() - t
    skip_oracle = os.environ.get("SKIP_ORACLE") == "1"
    out["cuda_moments"] = g.moments()
    os_ = None
    if not skip_oracle:
        t = time.time()
        o = H.Case(H.oracle_lib(), n=n, sc=sc)
        out["oracle_run"] = o.run(first, last)
        os_ = o.state()
        out["oracle_wall_s"] = time.time() - t
        out["cuda_vs_oracle"] = H.compare(gs, os_, 1e-6 if sc else 1e-9)
        out["oracle_moments"] = o.moments()
    if use_ref and os.path.exists(H.REF_DP):
        for dbl in (True, False):
            rs, log = H.run_reference(os.path.join(H.ROOT, "gpurun_out", "ref_work"), n, first, last, sc=sc, double=dbl)
            key = "ref_dp" if dbl else "ref_sp"
            out["cuda_vs_" + key] = H.compare(gs, rs, 1e-6 if sc else 1e-9)
            if os_ is not None:
                out["oracle_vs_" + key] = H.compare(os_, rs, 1e-6 if sc else 1e-9)
            out[key + "_stats"] = dict(zip(H.REF_STATS, rs["stats"].tolist()))
            out[key + "_log"] = [l for l in log.splitlines() if "time" in l or "wall" in l]
    from hpsim_b200 import HPSim
    print("JSON" + json.dumps(out, default=str))


if __name__ == "__main__":
    main()
