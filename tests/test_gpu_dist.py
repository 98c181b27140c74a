"""GPU: the row-partitioned hierarchy.  world_size 1 runs in-process; world_size 2 is launched with
torchrun when two GPUs are visible (skipped otherwise) — the same script the round's 2-GPU gpurun used."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _run(nproc, n, env=None, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py"), str(n)] + list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))


def test_dist_hierarchy_world1_matches_partitioned_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_check.py"), "14"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dist_hierarchy_world2_matches_partitioned_oracle():
    out = _run(2, 12)
    assert out.returncode == 0 and out.stdout.count("PASS") == 2, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_transport_failure_on_one_rank_falls_back_to_nccl_everywhere():
    """a rank that cannot export its window makes EVERY rank switch to the NCCL transport (collective detection);
    results still match the partitioned oracle"""
    out = _run(2, 10, env={"MLAMG_TEST_PEER_FAIL": "1"})
    assert out.returncode == 0 and out.stdout.count("PASS") == 2, out.stdout[-3000:] + out.stderr[-3000:]
    assert "peer transport unavailable" in out.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dist_hierarchy_world2_cube_geometry():
    """BASELINE config 5's layout at a reduced size: z-slabs of (2n) x (2n) x (n/4)"""
    out = _run(2, 16, extra=["--cube"])
    assert out.returncode == 0 and out.stdout.count("PASS") == 2, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
@pytest.mark.parametrize("extra", [[], ["--cube"]])
def test_dist_hierarchy_world4_interior_ranks(extra):
    """interior ranks with two neighbours exist from 3 ranks on"""
    out = _run(4, 12, extra=extra)
    assert out.returncode == 0 and out.stdout.count("PASS") == 4, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs")
def test_dist_hierarchy_world8():
    out = _run(8, 12, extra=["--cube"])
    assert out.returncode == 0 and out.stdout.count("PASS") == 8, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dist_hierarchy_world2_w32_interior_kernels():
    """the interior rows of the residual / prolongation passes on the W32 (warp-interleaved) copies, forced on at this size"""
    out = _run(2, 16, env={"MLAMG_W32_MIN_ROWS": "1"}, extra=["--cube"])
    assert out.returncode == 0 and out.stdout.count("PASS") == 2, out.stdout[-3000:] + out.stderr[-3000:]
