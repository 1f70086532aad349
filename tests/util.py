"""Shared helpers for the parity tests."""
import numpy as np


def stats_from_sums(sum_l, sum_l2, spp):
    mean = np.asarray(sum_l, float) / spp
    var = np.maximum(np.asarray(sum_l2, float) / spp - mean**2, 0.0) / spp  # pipelines/logic.py:957
    return mean, var


def z_scores(mean_a, var_a, mean_b, var_b, rel_floor=0.0):
    """Per-pixel paired z statistic (test_tools/regression.py:852-893).  `rel_floor` adds an
    fp32 arithmetic floor (relative) to the denominator for (nearly) deterministic pixels."""
    den = np.sqrt(var_a + var_b + (rel_floor * np.maximum(np.abs(mean_a), np.abs(mean_b))) ** 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        z = np.where(den > 0, (mean_a - mean_b) / den, np.where(np.isclose(mean_a, mean_b), 0.0, np.inf))
    return z


def sidak_ok(z, alpha=0.001):
    """Sidak-corrected acceptance of H0 for all pixels (regression.py:871-878)."""
    from scipy import stats

    z = np.atleast_1d(z)
    n = z.size
    a0 = 1.0 - (1.0 - alpha) ** (1.0 / n)
    zc = stats.norm.ppf(1.0 - a0 / 2.0)
    return bool(np.all(np.abs(z) <= zc)), zc
