#!/usr/bin/env python
"""GPU tool: where do the cycles of one tcgen05 conv launch go?  Runs single ops with the per-CTA cycle counters on
(instrumented build: `make -C interactive_deep_colorization_b200/csrc counters`, selected with
IDC_B200_LIB=interactive_deep_colorization_b200/lib/libidc_b200_counters.so) and prints the mean over CTAs.

    IDC_B200_LIB=... python tools/cta_counters.py [--n 64] [--opt name:val,...] [ops ...]
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--opt", default="")
ap.add_argument("ops", nargs="*", default=["c4_2", "c3_2", "c10_2", "up10", "c1_2"])
a = ap.parse_args()
opts = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in a.opt.split(",") if kv}
sd = synth.torch_state_dict(1234)
N = a.n
L, ab, m = synth.synthetic_batch(N, 256, seed=0)
ctx = util.make_ctx(sd, 256, 256, max_n=N, keep_conv10=True, use_graph=False, options=opts)
dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
ctx.forward_device(dL, dab, dm, 0.5)
torch.cuda.synchronize()
lib = ctx.lib
lib.idc_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib.idc_debug_counters(ctx.h, 1, None)
buf = np.zeros((148, 16), np.int64)
print("batch %d, options %r" % (N, opts))
print("op      mma_total wait_tempty wait_full | acc: wait_tfull drain epi(+splitK) | prologue first_data cta_life | splitK spin | kernel(us)   (kcycles, mean over CTAs with an MMA warp)")
for op in a.ops:
    ctx.run_op(op, N)
    torch.cuda.synchronize()
    lib.idc_debug_counters(ctx.h, 1, buf.ctypes.data)      # read + clear (warm-up launch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.run_op(op, N)
    e1.record()
    torch.cuda.synchronize()
    lib.idc_debug_counters(ctx.h, 1, buf.ctypes.data)
    act = buf[buf[:, 0] > 0]
    m_ = act.mean(0) / 1e3 if len(act) else buf.mean(0)
    print("%-6s %8.1f %10.1f %9.1f | %10.1f %7.1f %10.1f | %8.1f %9.1f %8.1f | %6.1f %5.1f | %8.1f   (%d CTAs)"
          % (op, m_[0], m_[1], m_[2], m_[3], m_[4], m_[5], m_[10], m_[6], m_[7], m_[8], m_[9], e0.elapsed_time(e1) * 1e3, len(act)))
ctx.close()
