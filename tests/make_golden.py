"""Generates tests/golden/*.npz from the reference's OWN CUDA code (oracle/_ref/ref_driver_dp, the
unmodified reference sources rebuilt for sm_100 with -DDOUBLE_PRECISION).  Run on a GPU box:
    gpurun -- 'python tests/make_golden.py && cp tests/golden/*.npz gpurun_out/'
The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and the CUDA product."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402

CASES = {
    # name: (n, first, last, sc)
    "lebt_sc_4k": (4096, "TBDB02", "spch_comp_7", True),
    "lebt_nosc_4k": (4096, "TBDB02", "spch_comp_7", False),
    "dtl_sc_4k": (4096, "TBDB02", "04QM30", True),
    "ccl_nosc_2k": (2048, "TBDB02", "06DT", False),
    # the whole linac (all 44 CCL modules; BASELINE C1 / C4 range)
    "full_nosc_2k": (2048, "TBDB02", "48DT", False),
    "full_sc_4k": (4096, "TBDB02", "48DT", True),
}


def main():
    out_dir = os.path.join(H.ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (n, first, last, sc) in CASES.items():
        st, log = H.run_reference(os.path.join(H.ROOT, "gpurun_out", "golden_work"), n, first, last, sc=sc, double=True)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), n=n, first=first, last=last, sc=sc,
                            stats=st["stats"], **{k: st[k] for k in ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")})
        print(name, "loss", int((st["loss"] != 0).sum()), "lloss", int((st["lloss"] != 0).sum()))
    # a beam file written by the reference's own Beam::PrintToFile, plus the same beam in binary
    work = os.path.join(H.ROOT, "gpurun_out", "golden_work")
    txt = os.path.join(out_dir, "ref_beam_print.txt")
    st, _ = H.run_reference_cfg(work, 256, ["gen dc", f"seed {H.SEED}", f"freq {H.FREQ}", f"ref_w {H.REF_W}", f"ref_phi {H.PHI_OFFSET!r}",
                                            f"shift_phi {H.PHI_OFFSET!r}"] + [f"{k} {v!r}" for k, v in H.DC.items()],
                                "TBDB02", "TBQL05V2", sc=False, extra=[f"print_to {txt}", "print_msg written by the reference"])
    np.savez_compressed(os.path.join(out_dir, "ref_beam_print.npz"), n=256, **{k: st[k] for k in ("x", "xp", "y", "yp", "phi", "w", "loss", "lloss")})


if __name__ == "__main__":
    main()
