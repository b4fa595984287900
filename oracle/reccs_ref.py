"""ORACLE (test infrastructure only -- never imported by the product path).

Colour suggestions at one pixel, SURVEY row f2.

* `sampled_reccs`  -- restatement of the reference procedure, /root/reference/data/colorize_image.py:322-354:
  cumulative sum of the pixel's pmf -> N uniform draws -> bin lookup (np.digitize) -> sklearn KMeans(K) on the
  drawn gamut points -> centres ordered by cluster occupancy.  Stochastic (np.random + KMeans init).
* `weighted_kmeans_pmf` -- the deterministic N -> infinity limit of it that the CUDA kernel
  (csrc/idc_heads.cu: ab_reccs_kernel) implements: weighted k-means over the gamut points with the pmf as
  weights; greedy farthest-point seeding (heaviest bin first, then argmax w * d^2, lowest index on ties),
  FP64 Lloyd iterations until the assignment is stable; best of n_init restarts.

Parity: the kernel is checked against `weighted_kmeans_pmf` to 1e-4 ab units; `weighted_kmeans_pmf` is checked
against `sampled_reccs` statistically (tests/test_reccs_cpu.py).  The reference holds no vectors for this
function (its output is random), so parity with the reference itself is statistical only.
"""
import numpy as np


def torch_gamut_points():
    """Bin -> ab table of the PyTorch wrapper (reference data/colorize_image.py:283; quirk q3)."""
    g = np.arange(-110, 120, 10)
    return np.array(np.meshgrid(g, g)).reshape((2, 529)).T.astype(np.float64)


def weighted_kmeans_pmf(pmf, pts, K, max_iter=100, n_init=8):
    """Best (lowest weighted inertia) of n_init restarts; restart v seeds from the bin of weight-rank v.
    Restarts within 1e-9 relative of the best count as ties -> lowest v."""
    runs = [_one_restart(pmf, pts, K, max_iter, v) for v in range(n_init)]
    e = np.array([weighted_inertia(pmf, pts, r[0]) for r in runs])
    pick = int(np.nonzero(e <= e.min() * (1.0 + 1e-9) + 1e-300)[0][0])
    return runs[pick]


def _one_restart(pmf, pts, K, max_iter, v):
    w = np.asarray(pmf, np.float64)
    w = w / w.sum()
    P = np.asarray(pts, np.float64)
    c = np.empty((K, 2))
    c[0] = P[np.lexsort((np.arange(w.size), -w))[v]]      # weight rank v, lowest index first among equals
    mind = ((P - c[0]) ** 2).sum(1)
    for j in range(1, K):
        c[j] = P[int(np.argmax(w * mind))]
        mind = np.minimum(mind, ((P - c[j]) ** 2).sum(1))
    labels = np.full(P.shape[0], -1)
    iters = 0
    while iters < max_iter:
        d = ((P[:, None, :] - c[None, :, :]) ** 2).sum(2)
        new = np.argmin(d, 1)
        if np.array_equal(new, labels):
            break
        labels = new
        for k in range(K):
            sel = labels == k
            m = w[sel].sum()
            if m > 0:
                c[k] = (w[sel, None] * P[sel]).sum(0) / m
        iters += 1
    mass = np.bincount(labels, weights=w, minlength=K)
    order = np.argsort(-mass, kind="stable")
    return c[order], mass[order], iters


def sampled_reccs(pmf, pts, K=5, N=25000, seed=0):
    from sklearn.cluster import KMeans
    rng = np.random.RandomState(seed)
    cmf = np.cumsum(np.asarray(pmf, np.float64))
    cmf = cmf / cmf[-1]
    inds = np.digitize(rng.uniform(0, 1.0, N), bins=cmf)
    samples = np.asarray(pts)[inds, :]
    km = KMeans(n_clusters=K, n_init=10, random_state=seed).fit(samples)
    cnt = np.histogram(km.labels_, np.arange(0, K + 1))[0]
    order = np.argsort(cnt)[::-1]
    return km.cluster_centers_[order, :], cnt[order] / float(N), float(km.inertia_) / N


def weighted_inertia(pmf, pts, centers):
    w = np.asarray(pmf, np.float64)
    w = w / w.sum()
    d = ((np.asarray(pts, np.float64)[:, None, :] - np.asarray(centers, np.float64)[None]) ** 2).sum(2)
    return float((w * d.min(1)).sum())


def synthetic_pmf(kind, seed=0):
    """Test pmfs over the 529 bins."""
    rng = np.random.RandomState(seed)
    P = torch_gamut_points()
    if kind == "uniform":
        return np.full(529, 1.0 / 529)
    if kind == "blobs":       # 3 Gaussian blobs of unequal mass + a small floor
        mu = np.array([[-60.0, 40.0], [50.0, 50.0], [20.0, -70.0]])
        a = np.array([0.6, 0.3, 0.1])
        p = sum(ai * np.exp(-((P - m) ** 2).sum(1) / (2 * 12.0 ** 2)) for ai, m in zip(a, mu))
        return p / p.sum() + 1e-6
    if kind == "softmax":     # what the dist head produces: softmax of random logits
        z = rng.randn(529) * 2.0
        e = np.exp(z - z.max())
        return e / e.sum()
    if kind == "peaked":      # nearly one-hot
        p = np.full(529, 1e-7)
        p[rng.randint(529)] = 1.0
        return p / p.sum()
    raise ValueError(kind)
