#!/usr/bin/env python
"""GPU tool: max|d ab| of the batch-1 interactive plan against the reference golden vector (tests/golden/lhn_256.npz,
5 random hints) under different plan-time options, plus the difference to the default plan.

    python tools/option_parity.py [name=opt:val,opt:val ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402


def main(specs):
    sd = synth.torch_state_dict(1234)
    g = util.golden("lhn_256.npz")
    L1 = g["img_l_mc"].astype(np.float32)[None]
    a1, m1 = synth.synthetic_hints(256, 5, 0)
    a1, m1 = a1[None].astype(np.float32), m1[None].astype(np.float32)
    ref = g["mc1_rand5_ab_raw"]
    base = None
    for spec in ["default="] + specs:
        name, _, body = spec.partition("=")
        opts = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in body.split(",") if kv}
        ctx = util.make_ctx(sd, 256, 256, max_n=1, dist=True, options=opts)
        r = ctx.forward_host(L1, a1, m1, 0.5, want_dist=True, want_rgb=True)
        r2 = ctx.forward_host(L1, a1, m1, 0.5, want_dist=True, want_rgb=True)
        if base is None:
            base = r
        print("[%s] max|d ab| vs reference golden %.3e   vs default plan %.3e   dist vs default %.3e   replay identical %s   launches %d"
              % (name, util.maxabs(r["ab"][0], ref), util.maxabs(r["ab"], base["ab"]), util.maxabs(r["dist"], base["dist"]),
                 bool(np.array_equal(r["ab"], r2["ab"])), ctx.last_launch_count()))
        ctx.close()


if __name__ == "__main__":
    main(sys.argv[1:])
