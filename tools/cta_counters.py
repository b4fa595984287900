#!/usr/bin/env python
"""GPU tool: where do the cycles of one tcgen05 conv launch go?  Runs ONE op (default c4_2) at batch 64
with the per-CTA cycle counters on and prints the mean over CTAs."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402

ops = sys.argv[1:] or ["c4_2", "c3_2", "c10_2", "up10", "c1_2"]
sd = synth.torch_state_dict(1234)
N = 64
L, ab, m = synth.synthetic_batch(N, 256, seed=0)
ctx = util.make_ctx(sd, 256, 256, max_n=N, keep_conv10=True, use_graph=False)
dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
ctx.forward_device(dL, dab, dm, 0.5)
torch.cuda.synchronize()
lib = ctx.lib
lib.idc_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib.idc_debug_counters(ctx.h, 1, None)
buf = np.zeros((148, 8), np.int64)
print("op      total  mma:wait_tempty wait_full | acc: wait_tfull drain epilogue   (kcycles, mean over CTAs)")
for op in ops:
    ctx.run_op(op, N)
    lib.idc_debug_counters(ctx.h, 1, buf.ctypes.data)
    m_ = buf.mean(0) / 1e3
    print("%-6s %7.1f  %9.1f %9.1f | %9.1f %7.1f %8.1f" % (op, m_[0], m_[1], m_[2], m_[3], m_[4], m_[5]))
ctx.close()
