"""Row f2 oracle checks (CPU): the deterministic weighted-k-means restatement vs the reference's
sample-and-cluster procedure (data/colorize_image.py:322-354), which is stochastic -> statistical parity."""
import numpy as np
import pytest

from oracle import reccs_ref as R
from tests import util

P = R.torch_gamut_points()


def _match(a, b):
    """greedy nearest matching distance between two centre sets"""
    b = list(map(tuple, b))
    worst = 0.0
    for c in a:
        d = [np.hypot(c[0] - q[0], c[1] - q[1]) for q in b]
        j = int(np.argmin(d))
        worst = max(worst, d[j])
        b.pop(j)
    return worst


def test_gamut_table_matches_reference_grid():
    assert P.shape == (529, 2) and tuple(P[1]) == (-100.0, -110.0) and tuple(P[23]) == (-110.0, -100.0)   # quirk q3


@pytest.mark.parametrize("kind,seed", [("blobs", 0), ("softmax", 1), ("softmax", 2), ("uniform", 0)])
def test_weighted_limit_is_at_least_as_good_as_sampling(kind, seed):
    pmf = R.synthetic_pmf(kind, seed)
    c, mass, iters = R.weighted_kmeans_pmf(pmf, P, 5)
    cs, confs, _ = R.sampled_reccs(pmf, P, 5, 25000, seed)
    assert abs(mass.sum() - 1) < 1e-12 and np.all(np.diff(mass) <= 1e-15)
    # objective of the population problem the reference approximates by sampling
    assert R.weighted_inertia(pmf, P, c) <= 1.01 * R.weighted_inertia(pmf, P, cs)


def test_well_separated_modes_agree_with_sampling():
    pmf = np.full(529, 1e-9)
    idx, w = [30, 262, 500], [0.55, 0.3, 0.15]
    for i, wi in zip(idx, w):
        pmf[i] = wi
    c, mass, _ = R.weighted_kmeans_pmf(pmf, P, 3)
    cs, confs, _ = R.sampled_reccs(pmf, P, 3, 25000, 0)
    assert _match(c, cs) < 0.5 and np.allclose(mass, confs, atol=0.01)
    assert np.allclose(c, P[idx], atol=1e-3) and np.allclose(mass, w, atol=1e-6)


def test_deterministic_and_restart_monotone():
    pmf = R.synthetic_pmf("softmax", 3)
    a = R.weighted_kmeans_pmf(pmf, P, 7, n_init=8)
    b = R.weighted_kmeans_pmf(pmf, P, 7, n_init=8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    e1 = R.weighted_inertia(pmf, P, R.weighted_kmeans_pmf(pmf, P, 7, n_init=1)[0])
    assert R.weighted_inertia(pmf, P, a[0]) <= e1 + 1e-12


def test_against_reference_run_in_golden():
    """tests/golden/lhn_dist_256.npz holds get_ab_reccs(128,128,K=9) of the reference itself (seeded)."""
    gd = util.golden("lhn_dist_256.npz")
    pmf = gd["dist_rows"][:, 4, 4]                      # dist[:, 128, 128] = 64-grid (32,32) = rows[::8] index 4
    ref = gd["reccs_128_128_K9"]
    c, mass, _ = R.weighted_kmeans_pmf(pmf, P, 9)
    assert R.weighted_inertia(pmf, P, c) <= 1.01 * R.weighted_inertia(pmf, P, ref)
