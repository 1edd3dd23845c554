"""Parity at the shapes the bench and the multi_ref workload actually run: the production
U-Net at B = 10 (five view groups per call: the bench's plan), with 1 and with 4 reference views, at the first, a
middle and the last DDIM timestep - each held to 1e-2 on the noise prediction, with the per-block errors printed.
The full 3-seed matrix is `python tests/parity_budget.py` (table in DESIGN.md)."""
import pytest
import torch

from tests.parity_budget import run_matrix

pytestmark = pytest.mark.gpu
EPS_TOL = 1e-2


def test_production_parity_at_bench_shapes(cuda_device):
    rows = run_matrix(cuda_device, seeds=(0,), timesteps=(1, 501, 991), Rs=(1, 4), B=10)
    assert len(rows) == 6
    for r in rows:
        assert r["final"] < EPS_TOL, r
