"""CPU restatement of the Caffe-only pieces of the hot path (no Caffe runtime exists here and
the reference holds no test vectors for them: PARITY UNPINNED -- see oracle/__init__.py).

  global_hints_vector   models/global_model/deploy_nodist.prototxt:38-172
                        glob_conv1(glob_ab_313_mask[314]) + s_conv1(s_avg_mask[2]) -> ReLU -> BatchNorm,
                        then 3 x (1x1 conv 512->512, ReLU, BatchNorm); the result is broadcast over
                        space (SpatialRepLayer, caffe_files/caffe_traininglayers.py:14-50) and added to
                        conv4_3norm (:501-527) -- that add is `glob_add` in oracle/lhn_ref.py.
                        Input driver: data/colorize_image.py:452-463.

Test infrastructure only.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def synthetic_glob_state_dict(seed=4321):
    """Keys glob.{0..3}.weight [512, 316|512], .bias, .bn.{weight,bias,running_mean,running_var}.
    Layer 0's weight is [glob_conv1 (314 cols) | s_conv1 (2 cols)], bias = sum of both biases."""
    rng = np.random.RandomState(seed)
    sd = {}
    for l in range(4):
        cin = 316 if l == 0 else 512
        std = np.sqrt(2.0 / (8.0 if l == 0 else cin))     # layer-0 inputs are a pmf + indicators: small fan-in
        sd["glob.%d.weight" % l] = (rng.standard_normal((512, cin)) * std).astype(np.float32)
        sd["glob.%d.bias" % l] = rng.uniform(-0.1, 0.1, 512).astype(np.float32)
        sd["glob.%d.bn.weight" % l] = rng.uniform(0.6, 1.4, 512).astype(np.float32)
        sd["glob.%d.bn.bias" % l] = (rng.standard_normal(512) * 0.1).astype(np.float32)
        sd["glob.%d.bn.running_mean" % l] = rng.uniform(0.1, 0.6, 512).astype(np.float32)
        sd["glob.%d.bn.running_var" % l] = rng.uniform(0.3, 1.2, 512).astype(np.float32)
    return sd


def global_hints_vector(gsd, glob316):
    """glob316 [N,316] = [313 histogram, indicator, mean saturation, indicator] -> [N,512]."""
    x = torch.as_tensor(np.asarray(glob316), dtype=torch.float32)
    for l in range(4):
        t = lambda k: torch.as_tensor(np.asarray(gsd["glob.%d.%s" % (l, k)]), dtype=torch.float32)
        x = F.relu(F.linear(x, t("weight"), t("bias")))
        x = F.batch_norm(x, t("bn.running_mean"), t("bn.running_var"), t("bn.weight"), t("bn.bias"), False, 0.0, BN_EPS)
    return x
