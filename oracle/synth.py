"""Deterministic synthetic weights and inputs (numpy RandomState => bit-stable
across machines).  There are no real weights offline (the reference fetches them
with wget, /root/reference/models/fetch_models.sh:2-7), so every parity check runs
on these.  Key names / shapes follow the reference state_dict
(/root/reference/models/pytorch/model.py:5-132).

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np

# (key, kind, cin, cout, k) in reference order.  conv weights are OIHW, deconv IOHW.
CONV_LAYERS = [
    ("model1.0", "conv", 4, 64, 3), ("model1.2", "conv", 64, 64, 3),
    ("model2.0", "conv", 64, 128, 3), ("model2.2", "conv", 128, 128, 3),
    ("model3.0", "conv", 128, 256, 3), ("model3.2", "conv", 256, 256, 3), ("model3.4", "conv", 256, 256, 3),
    ("model4.0", "conv", 256, 512, 3), ("model4.2", "conv", 512, 512, 3), ("model4.4", "conv", 512, 512, 3),
    ("model5.0", "conv", 512, 512, 3), ("model5.2", "conv", 512, 512, 3), ("model5.4", "conv", 512, 512, 3),
    ("model6.0", "conv", 512, 512, 3), ("model6.2", "conv", 512, 512, 3), ("model6.4", "conv", 512, 512, 3),
    ("model7.0", "conv", 512, 512, 3), ("model7.2", "conv", 512, 512, 3), ("model7.4", "conv", 512, 512, 3),
    ("model8up.0", "deconv", 512, 256, 4), ("model3short8.0", "conv", 256, 256, 3),
    ("model8.1", "conv", 256, 256, 3), ("model8.3", "conv", 256, 256, 3),
    ("model9up.0", "deconv", 256, 128, 4), ("model2short9.0", "conv", 128, 128, 3),
    ("model9.1", "conv", 128, 128, 3),
    ("model10up.0", "deconv", 128, 128, 4), ("model1short10.0", "conv", 64, 128, 3),
    ("model10.1", "conv", 128, 128, 3),
    ("model_class.0", "conv", 256, 529, 1),
    ("model_out.0", "conv", 128, 2, 1),
]
BN_LAYERS = [("model1.4", 64), ("model2.4", 128), ("model3.6", 256), ("model4.6", 512),
             ("model5.6", 512), ("model6.6", 512), ("model7.6", 512), ("model8.5", 256),
             ("model9.3", 128)]


def synthetic_state_dict(seed=1234):
    """Seeded Kaiming-fan-in convs, randomised BN statistics, small regression-head
    gain so tanh stays unsaturated.  Returns {key: float32 ndarray} (plus int64
    num_batches_tracked) with exactly the reference's state_dict keys."""
    rng = np.random.RandomState(seed)
    sd = {}
    for key, kind, cin, cout, k in CONV_LAYERS:
        if kind == "conv":
            fan_in = cin * k * k
            shape = (cout, cin, k, k)
        else:  # ConvTranspose2d 4x4 s2: every output pixel sees 2x2 taps
            fan_in = cin * 4
            shape = (cin, cout, k, k)
        std = np.sqrt(2.0 / fan_in)
        if key == "model_out.0":
            std = 0.25 / np.sqrt(cin)
        if key == "model_class.0":
            std = 12.0 / np.sqrt(cin)      # logits*0.2 should have O(1) spread
        if "short" in key or "up" in key:
            std *= np.sqrt(0.5)            # the two branches are summed
        sd[key + ".weight"] = (rng.standard_normal(shape) * std).astype(np.float32)
        sd[key + ".bias"] = (rng.uniform(-0.1, 0.1, cout)).astype(np.float32)
    for key, c in BN_LAYERS:
        sd[key + ".weight"] = rng.uniform(0.6, 1.4, c).astype(np.float32)
        sd[key + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[key + ".running_mean"] = rng.uniform(0.1, 0.6, c).astype(np.float32)
        sd[key + ".running_var"] = rng.uniform(0.3, 1.2, c).astype(np.float32)
        sd[key + ".num_batches_tracked"] = np.array(1000, dtype=np.int64)
    return sd


def torch_state_dict(seed=1234):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in synthetic_state_dict(seed).items()}


def put_point(input_ab, mask, loc, p, val):
    """Notebook helper, /root/reference/DemoInteractiveColorization.ipynb cell 5 (:131-139)."""
    input_ab[:, loc[0] - p:loc[0] + p + 1, loc[1] - p:loc[1] + p + 1] = np.array(val)[:, np.newaxis, np.newaxis]
    mask[:, loc[0] - p:loc[0] + p + 1, loc[1] - p:loc[1] + p + 1] = 1
    return (input_ab, mask)


def synthetic_hints(X, nhints, seed):
    """centres randint(8, X-8), p=3 (7x7), val U(-80,80)."""
    rng = np.random.RandomState(seed)
    ab = np.zeros((2, X, X), dtype=np.float64)
    mask = np.zeros((1, X, X), dtype=np.float64)
    if nhints > 0:
        centres = rng.randint(8, X - 8, (nhints, 2))
        for i in range(nhints):
            val = rng.uniform(-80, 80, 2)
            put_point(ab, mask, centres[i], 3, val)
    return ab, mask


def synthetic_batch(N, X, seed=0, max_hints=10):
    """L = rand*100 (then -50), per-image 0..max_hints hints,
    seed = base seed + image index.  Returns float32 (L_mc[N,1,X,X], ab[N,2,X,X], mask[N,1,X,X])."""
    L = np.empty((N, 1, X, X), np.float32)
    ab = np.empty((N, 2, X, X), np.float32)
    mask = np.empty((N, 1, X, X), np.float32)
    for i in range(N):
        rng = np.random.RandomState(seed + i)
        L[i, 0] = (rng.rand(X, X) * 100.0 - 50.0).astype(np.float32)
        nh = int(rng.randint(0, max_hints + 1))
        a, m = synthetic_hints(X, nh, seed + i + 7919)
        ab[i] = a
        mask[i] = m
    return L, ab, mask


def synthetic_glob(N, seed=0):
    """Dirichlet(0.1) 313-bin histogram || indicator 1,
    s_avg in U[0,1] || indicator 1."""
    rng = np.random.RandomState(seed)
    hist = rng.dirichlet(np.full(313, 0.1), N).astype(np.float32)
    glob_ab = np.concatenate([hist, np.ones((N, 1), np.float32)], axis=1)
    sat = np.concatenate([rng.uniform(0, 1, (N, 1)).astype(np.float32), np.ones((N, 1), np.float32)], axis=1)
    return glob_ab, sat
