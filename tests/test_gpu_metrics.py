"""GPU: device-side evaluation metrics (poseidon_b200.metrics, scot_lp_plane_sums) vs the numpy restatement of
scOT/metrics.py (oracle/metrics_oracle.py). fp32 tree sums vs fp64 numpy: 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("shape", [(7, 5, 128, 128), (3, 1, 64, 64), (2, 4, 33, 17)])
def test_relative_lp_error_matches_reference_metrics(shape, p):
    from poseidon_b200 import metrics as M

    g = torch.Generator().manual_seed(5)
    pred, y = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    y[0, 0] = 0.0  # a zero plane: exercises the zero-normaliser guard only when the whole sample is zero
    ref = MO.relative_lp_error(pred.double().numpy(), y.double().numpy(), p=p)
    got = M.relative_lp_error(pred.cuda(), y.cuda(), p=p)
    assert np.allclose(got, ref, rtol=1e-5)
    assert np.allclose(M.mean_relative_lp_error(pred.cuda(), y.cuda(), p=p), ref.mean(), rtol=1e-5)
    assert np.allclose(M.median_relative_lp_error(pred.cuda(), y.cuda(), p=p), np.median(ref), rtol=1e-5)
    # all-zero target -> normaliser 1e-10 (metrics.py:27-30)
    z = torch.zeros(shape)
    assert np.allclose(M.relative_lp_error(pred.cuda(), z.cuda(), p=p), MO.relative_lp_error(pred.double().numpy(), z.double().numpy(), p=p), rtol=1e-5)


def test_error_statistics_match_compute_metrics():
    from poseidon_b200 import metrics as M

    g = torch.Generator().manual_seed(6)
    pred, y = torch.randn(9, 5, 64, 64, generator=g), torch.randn(9, 5, 64, 64, generator=g)
    sl = [0, 1, 3, 4, 5]
    ref = MO.group_statistics(pred.double().numpy(), y.double().numpy(), sl)
    got = M.error_statistics(pred.cuda(), y.cuda(), sl, ["rho", "uv", "p", "tr"])
    for name, r in zip(["rho", "uv", "p", "tr"], ref):
        for k, v in r.items():
            # errors are ~100 (percent); fp32 plane sums carry ~1e-7 relative error, which the std (a difference of
            # nearly equal numbers) sees in absolute terms
            assert abs(got[name + "/" + k] - v) < 1e-5 * abs(v) + 1e-4, (name, k)
    assert abs(got["mean_relative_l1_error"] - np.mean([r["mean_relative_l1_error"] for r in ref])) < 1e-4
    assert abs(got["mean_over_median_relative_l1_error"] - np.mean([r["median_relative_l1_error"] for r in ref])) < 1e-4
