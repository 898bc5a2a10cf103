"""N > 1 path on CPU: two gloo ranks share one beam through the sharded CPU oracle (see
tests/sharded_oracle_worker.py) and must agree with the single-process oracle.  The same ownership
(interleaved), the same exchange points (segment epilogue, lloss pass, moments, charge mesh) and the
same global-index quirks (H1, H2) as the CUDA product; the GPU counterpart is tests/sharded_check.py."""
import json
import os
import subprocess
import sys

import pytest

import harness as H

WORKER = os.path.join(H.ROOT, "tests", "sharded_oracle_worker.py")


def run_ranks(world, n, first, last, sc, port):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, str(n), first, last, "1" if sc else "0"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("SHARDED_ORACLE ")]
    assert line, p.stdout[-2000:]
    return json.loads(line[-1][len("SHARDED_ORACLE "):])


@pytest.mark.parametrize("world,n,last,sc", [(2, 4096, "spch_comp_7", True), (2, 1001, "01QM04", True), (3, 2048, "TDDB01", False)])
def test_sharded_oracle_matches_single_process(world, n, last, sc):
    H.oracle_lib()  # build once, before the ranks race for it
    rep = run_ranks(world, n, "TBDB02", last, sc, 29620 + world + n % 7)
    assert rep["loss_equal"] and rep["lloss_equal"], rep
    assert rep["counts"][0] == rep["counts"][1], rep
    assert rep["kicks"][0] == rep["kicks"][1] and rep["remesh"][0] == rep["remesh"][1], rep
    assert rep["worst"] < 1e-9, rep                       # only the summation order differs
    assert max(rep["moments_rel"].values()) < 1e-9, rep
    if sc:
        assert rep["allreduce_calls"] > 5 * rep["kicks"][0], rep  # the kick chain really crosses ranks
