#!/usr/bin/env python
"""GPU tool for ncu: a few batch-N forwards of one ctx (args: N [forwards])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sd = synth.torch_state_dict(1234)
L, ab, m = synth.synthetic_batch(n, 256, seed=0)
ctx = util.make_ctx(sd, 256, 256, max_n=n, dist=True, use_graph=False)
dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
for _ in range(reps):
    ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True)
torch.cuda.synchronize()
ctx.close()
