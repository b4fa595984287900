#!/usr/bin/env python
"""GPU tool: per-op device time of the batch-1 interactive path (max_n=1 ctx, dist head on) and the
wall-clock split of idc_forward_host (H2D / graph / D2H)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402


def run(tag):
    sd = synth.torch_state_dict(1234)
    L, ab, m = synth.synthetic_batch(1, 256, seed=0)
    ctx = util.make_ctx(sd, 256, 256, max_n=1, dist=True, use_graph=False)
    dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
    for _ in range(5):
        ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True)
    torch.cuda.synchronize()
    ctx.set_profiling(True)
    for _ in range(20):
        ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True)
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    tot = sum(ms for _, ms, _ in prof)
    print("[%s] per-op us (sum %.1f us): " % (tag, tot * 1e3) + " ".join("%s=%.1f" % (n, ms * 1e3) for n, ms, _ in prof))
    ctx.close()
    for want_dist in (True, False):
        ctx = util.make_ctx(sd, 256, 256, max_n=1, dist=True, use_graph=True)
        ts = []
        for i in range(30):
            t = time.perf_counter()
            ctx.forward_host(L, ab, m, 0.5, want_dist=want_dist, want_rgb=True)
            ts.append((time.perf_counter() - t) * 1e3)
        print("[%s] forward_host graph want_dist=%s p50 %.3f ms  min %.3f ms" % (tag, want_dist, np.percentile(ts[5:], 50), min(ts)))
        ctx.close()


if __name__ == "__main__":
    for sk in ("1", None):
        if sk:
            os.environ["IDC_SPLIT_K"] = sk
        elif "IDC_SPLIT_K" in os.environ:
            del os.environ["IDC_SPLIT_K"]
        run("split_k=%s" % (sk or "auto"))
