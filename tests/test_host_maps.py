"""CPU check of the kernel element maps: hpsim_b200/csrc/hps_kernels.cuh compiled for the HOST
(tests/host_maps.cu).  The fast maps the product launches by default (reciprocals of block-uniform
quantities, rsqrt instead of sqrt-then-divide) must agree with the reference-order maps
(simulation_kernel.cu:16-871, space_charge_kernel.cu:573-706) at rounding level on every op record of
the full linac, with identical loss decisions."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_fast_maps_match_reference_order_maps():
    import harness as H
    exe = H.build_tool("host_maps")
    # TBDB02 .. 48DT: every drift, quad and RF-gap record of the LANSCE linac, 16 random particles each
    p = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "lansce.lat"), "62", "6537", "16", "7"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    m = re.search(r"WORST ([0-9.eE+-]+) LOSS_MISMATCH (\d+)", p.stdout)
    assert m, p.stdout
    assert float(m.group(1)) < 1e-11, p.stdout
    assert int(m.group(2)) == 0, p.stdout
    for name in ("drift", "quad_half", "gap_first", "gap_second", "sc_kick"):
        assert re.search(rf"^{name}\s+ops\s+\d+", p.stdout, re.M), p.stdout
