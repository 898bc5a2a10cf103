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


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
@pytest.mark.parametrize("first,last,kicks,gaps", [(62, 6537, 8894, 5124), (62, 129, 236, 0), (130, 478, 406, None), (62, 710, 967, None)])
def test_planner_counts(first, last, kicks, gaps):
    """The host planner (hps_plan.cpp) on the LANSCE lattice at interval 0.025 m: the kick counts SURVEY.md 8d derives
    from the reference's element loop (drifts ceil(L / interval) kicks, every quad / gap / dipole one), the RF gaps of
    the full linac, one segment per kick plus the sweeps that end at an "in" aperture or at the end of the range, and
    the state the engine leaves behind (805 MHz and the 800 MeV design energy after the last CCL gap)."""
    import harness as H
    exe = H.build_tool("host_maps")
    p = subprocess.run([exe, "plan", os.path.join(ROOT, "tests", "golden", "lansce.lat"), str(first), str(last), "0.025"],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    m = dict(re.findall(r"(\w+) ([0-9.]+)", p.stdout))
    assert int(m["kicks"]) == kicks == int(m["kick_segments"]) == int(m["kick_ops"]), p.stdout
    if gaps is not None:
        assert int(m["gaps"]) == gaps, p.stdout
    assert int(m["segments"]) == kicks + 1 + int(m["aperture_syncs"]), p.stdout
    assert int(m["max_ops_per_segment"]) <= 24
    if last == 6537:
        assert abs(float(m["freq"]) - 805.0) < 1e-9 and 795.0 < float(m["design_w"]) < 805.0, p.stdout


def test_pattern_sweeps_cover_the_full_linac():
    """The ten op sequences k_segment is compiled for (SEQ_* in hps_kernels.cu) against the planner's own segments of the full
    linac: together they are at least 99.5 % of the sweeps (9 086 of 9 108), led by the CCL cell and the drift step."""
    import harness as H
    exe = H.build_tool("host_maps")
    p = subprocess.run([exe, "plan", os.path.join(ROOT, "tests", "golden", "lansce.lat"), "62", "6537", "0.025"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, HPS_PLAN_PATTERNS="1"))
    assert p.returncode == 0, p.stderr
    hist = {m.group(2).strip(): int(m.group(1)) for m in re.finditer(r"PATTERN\s+(\d+)\s+(.*)", p.stdout)}
    compiled = ["KICK gap2c gap1c | kick", "KICK drift drift | kick", "KICK drift quad | kick", "KICK quad drift | kick",
                "KICK quad gap1d | kick", "KICK gap2d quad | kick", "capert drift | kick", "KICK drift | aperture",
                "KICK gap2c | aperture", "capert gap1c | kick"]
    total = sum(hist.values())
    covered = sum(hist.get(k, 0) for k in compiled)
    assert total == 9108, total
    assert hist["KICK gap2c gap1c | kick"] == 4855 and hist["KICK drift drift | kick"] == 3030, hist
    assert covered >= 0.995 * total, (covered, total)
