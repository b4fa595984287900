#!/usr/bin/env python
"""GPU tool: conv1_2 (64 -> 64, 256^2) with the halo-tile A operand + resident weights (option halo:3 adds it to the
default plan) vs the per-tap MT=2 pair kernel: per-op time at batch 64 and max|d| of conv1_2 / the final ab map."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = synth.torch_state_dict(1234)
L, ab, m = synth.synthetic_batch(n, 256, seed=0)
dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
res = {}
for tag, opts in (("per-tap", {}), ("halo+resident", {"halo": 3})):
    ctx = util.make_ctx(sd, 256, 256, max_n=n, use_graph=False, keep_conv10=False, options=opts)
    for _ in range(3):
        out = ctx.forward_device(dL, dab, dm, 0.5)
    torch.cuda.synchronize()
    ctx.set_profiling(True)
    for _ in range(10):
        ctx.forward_device(dL, dab, dm, 0.5)
    prof = dict((nm, ms) for nm, ms, _ in ctx.get_profile())
    ctx.set_profiling(False)
    act = ctx.get_activation("conv1_2", min(n, 4)).cpu().numpy()
    res[tag] = (act, out["ab"].cpu().numpy())
    print("batch %d %-14s c1_2 %.1f us   pack+conv1_1 %.1f   c2_1 %.1f   forward sum %.2f ms"
          % (n, tag, prof["c1_2"] * 1e3, prof["pack+conv1_1"] * 1e3, prof["c2_1"] * 1e3, sum(prof.values())))
    ctx.close()
print("max|d conv1_2| = %.3e (max|conv1_2| %.2f)   max|d ab| = %.3e"
      % (float(np.abs(res["per-tap"][0] - res["halo+resident"][0]).max()), float(np.abs(res["per-tap"][0]).max()),
         float(np.abs(res["per-tap"][1] - res["halo+resident"][1]).max())))
