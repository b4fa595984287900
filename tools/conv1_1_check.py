#!/usr/bin/env python
"""GPU tool: model1.0 (conv1_1) on the tensor cores vs the exact FP32 CUDA-core kernel: max|d a1_1| over a few
geometries (partial tiles, tiny images, batches) and the time of the `pack+conv1_1` slot at batch 1 and batch 64."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402

sd = synth.torch_state_dict(1234)
for (H, W, n) in ((8, 8, 1), (24, 40, 3), (64, 64, 2), (256, 256, 1), (256, 256, 4)):
    rs = np.random.RandomState(H + n)
    L = rs.uniform(-50, 50, (n, 1, H, W)).astype(np.float32)
    ab = (rs.uniform(-100, 100, (n, 2, H, W)) * (rs.rand(n, 1, H, W) < 0.2)).astype(np.float32)
    m = (rs.rand(n, 1, H, W) < 0.2).astype(np.float32)
    acts = {}
    for mode in (1, 0):
        ctx = util.make_ctx(sd, H, W, max_n=n, options={"conv1_1_umma": mode})
        ctx.forward_device(util.dev(L), util.dev(ab), util.dev(m), 0.5)
        torch.cuda.synchronize()
        acts[mode] = ctx.get_activation("a1_1", n).cpu().numpy()
        ctx.close()
    ref = acts[0]
    print("a1_1 %dx%d n=%d: max|tensor - fp32| = %.3e  (max|a1_1| = %.2f, mean %.3f)"
          % (H, W, n, float(np.abs(acts[1] - ref).max()), float(np.abs(ref).max()), float(np.abs(ref).mean())))
for n in (1, 64):
    L, ab, m = synth.synthetic_batch(n, 256, seed=0)
    dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
    for mode in (1, 0):
        ctx = util.make_ctx(sd, 256, 256, max_n=n, use_graph=False, options={"conv1_1_umma": mode})
        for _ in range(3):
            ctx.forward_device(dL, dab, dm, 0.5)
        torch.cuda.synchronize()
        ctx.set_profiling(True)
        for _ in range(10):
            ctx.forward_device(dL, dab, dm, 0.5)
        prof = ctx.get_profile()
        ctx.set_profiling(False)
        print("batch %d conv1_1_umma=%d: pack+conv1_1 %.1f us, c1_2 %.1f us, forward sum %.1f us"
              % (n, mode, prof[0][1] * 1e3, prof[1][1] * 1e3, sum(p[1] for p in prof) * 1e3))
        ctx.close()
