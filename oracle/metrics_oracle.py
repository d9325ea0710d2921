"""CPU restatement (numpy) of the reference's evaluation metrics — TEST INFRASTRUCTURE, never imported by the product.

Follows /root/reference/scOT/metrics.py:12-36 (`relative_lp_error`) and the statistics of the `compute_metrics`
closure in /root/reference/scOT/train.py:344-398. Pinned in tests/test_oracle_golden.py::test_metrics_oracle_known_answer
by closed-form known-answer cases (the reference ships no fixtures for it)."""
import numpy as np


def relative_lp_error(preds, targets, p=1, return_percent=True):
    n, c = preds.shape[:2]
    a = preds.reshape(n, c, -1)
    b = targets.reshape(n, c, -1)
    err = np.sum(np.abs(a - b) ** p, axis=-1)
    norm = np.sum(np.sum(np.abs(b) ** p, axis=-1), axis=-1)
    norm = np.where(norm == 0, 1e-10, norm)
    e = (np.sum(err, axis=-1) / norm) ** (1 / p)
    return e * 100 if return_percent else e


def group_statistics(preds, targets, channel_slice_list):
    out = []
    for i in range(len(channel_slice_list) - 1):
        lo, hi = channel_slice_list[i], channel_slice_list[i + 1]
        e = relative_lp_error(preds[:, lo:hi], targets[:, lo:hi], p=1, return_percent=True)
        out.append({"median_relative_l1_error": np.median(e, axis=0), "mean_relative_l1_error": np.mean(e, axis=0),
                    "std_relative_l1_error": np.std(e, axis=0), "min_relative_l1_error": np.min(e, axis=0),
                    "max_relative_l1_error": np.max(e, axis=0)})
    return out
