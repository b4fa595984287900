#!/usr/bin/env python
"""GPU tool: where does a click (BASELINE config 5) spend its time?

  * wall clock of idc_forward_host with the page-locked click buffers, and of idc_fetch_dist on its own
  * device span of the click graph (two events around the graph launch: copy nodes + kernels)
  * kernels only: a torch-captured graph of idc_forward (device-resident inputs) replayed back to back
  * plain pinned copies of the click's sizes (H2D 1 MB, D2H 0.7 MB) for scale
  * the SM clock NVML reports right after each click

    python tools/click_breakdown.py [opt:val,opt:val ...]
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from interactive_deep_colorization_b200 import _lib  # noqa: E402
from interactive_deep_colorization_b200 import colorize_image as CI  # noqa: E402
from oracle import synth  # noqa: E402
from tests import util  # noqa: E402


def sm_clock_reader():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception as e:  # noqa: BLE001
        print("NVML not available:", e)
        return lambda: -1


def pct(v, q):
    return float(np.percentile(v, q))


def run(tag, options, X=256):
    lib = _lib.load()
    lib.idc_debug_graph_timing.restype = ctypes.c_int
    lib.idc_debug_graph_timing.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    clk = sm_clock_reader()
    sd = synth.torch_state_dict(1234)
    L, ab, m = synth.synthetic_batch(1, X, seed=0)
    rs = np.random.RandomState(0)
    ctx = util.make_ctx(sd, X, X, max_n=1, dist=True, use_graph=True, options=options)
    ctx.set_dist_resident(True)
    buf = ctx.click_buffers(1)
    buf["L_mc"][...] = L; buf["ab"][...] = 0; buf["mask"][...] = 0
    lib.idc_debug_graph_timing(ctx.h, 1, None)
    t_fwd, t_fetch, t_graph, clocks = [], [], [], []
    ms = ctypes.c_float(0)
    for i in range(45):
        loc = rs.randint(8, X - 8, 2)
        CI.put_point(buf["ab"][0], buf["mask"][0], loc, 3, rs.uniform(-80, 80, 2))
        t0 = time.perf_counter()
        ctx.forward_host(buf["L_mc"], buf["ab"], buf["mask"], 0.5, want_rgb=True, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"])
        t1 = time.perf_counter()
        ctx.fetch_dist(0, int(loc[0]) // 4, int(loc[1]) // 4)
        t2 = time.perf_counter()
        clocks.append(clk())
        lib.idc_debug_graph_timing(ctx.h, 1, ctypes.byref(ms))
        t_fwd.append((t1 - t0) * 1e3); t_fetch.append((t2 - t1) * 1e3); t_graph.append(ms.value)
    t_fwd, t_fetch, t_graph, clocks = t_fwd[5:], t_fetch[5:], t_graph[5:], clocks[5:]
    print("[%s] click: forward_host wall p50 %.3f ms (min %.3f)  graph device span p50 %.3f ms (min %.3f)  fetch_dist p50 %.3f ms"
          % (tag, pct(t_fwd, 50), min(t_fwd), pct(t_graph, 50), min(t_graph), pct(t_fetch, 50)))
    print("[%s] click: total p50 %.3f ms; SM clock after each click: median %d MHz (min %d, max %d)"
          % (tag, pct(np.add(t_fwd, t_fetch), 50), int(np.median(clocks)), min(clocks), max(clocks)))
    # the shipped click: resident image (idc_set_image), announced click (idc_set_click, K = 9 suggestions on the side branch)
    ctx.set_image(buf["L_mc"])
    t_all, t_graph2 = [], []
    for i in range(45):
        loc = rs.randint(8, X - 8, 2)
        CI.put_point(buf["ab"][0], buf["mask"][0], loc, 3, rs.uniform(-80, 80, 2))
        y4, x4 = int(loc[0]) // 4, int(loc[1]) // 4
        t0 = time.perf_counter()
        ctx.set_click(0, y4, x4, 9)
        ctx.forward_host(None, buf["ab"], buf["mask"], 0.5, want_rgb=True, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"])
        ctx.fetch_dist(0, y4, x4)
        ctx.ab_reccs(0, y4, x4, K=9)
        t_all.append((time.perf_counter() - t0) * 1e3)
        lib.idc_debug_graph_timing(ctx.h, 1, ctypes.byref(ms))
        t_graph2.append(ms.value)
    print("[%s] announced click, resident image (forward + pmf + 9 suggestions): wall p50 %.3f ms (min %.3f), graph span p50 %.3f ms"
          % (tag, pct(t_all[5:], 50), min(t_all[5:]), pct(t_graph2[5:], 50)))
    ctx.set_click(0, -1, 0, 0)
    t_all = []
    for i in range(30):
        t0 = time.perf_counter()
        ctx.forward_host(None, buf["ab"], buf["mask"], 0.5, want_rgb=True, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"])
        t_all.append((time.perf_counter() - t0) * 1e3)
        lib.idc_debug_graph_timing(ctx.h, 1, ctypes.byref(ms))
        t_graph2.append(ms.value)
    print("[%s] resident image, click mode off: forward_host wall p50 %.3f ms, graph span p50 %.3f ms"
          % (tag, pct(t_all[5:], 50), pct(t_graph2[-20:], 50)))
    # spaced clicks: one per 50 ms, as a user would issue them (does the clock drop between clicks?)
    t_sp, c_sp = [], []
    for i in range(12):
        time.sleep(0.05)
        t0 = time.perf_counter()
        ctx.forward_host(buf["L_mc"], buf["ab"], buf["mask"], 0.5, want_rgb=True, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"])
        t_sp.append((time.perf_counter() - t0) * 1e3)
        c_sp.append(clk())
    print("[%s] clicks 50 ms apart: forward_host wall p50 %.3f ms, SM clock median %d MHz" % (tag, pct(t_sp[2:], 50), int(np.median(c_sp))))
    lib.idc_debug_graph_timing(ctx.h, 0, None)
    ctx.close()

    # kernels only: device-resident inputs, torch graph of idc_forward, replayed back to back
    ctx = util.make_ctx(sd, X, X, max_n=1, dist=True, use_graph=False, options=options)
    dL, dab, dm = util.dev(L), util.dev(ab), util.dev(m)
    out = ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True, out_ab=out["ab"], out_dist=out["dist"], out_rgb=out["rgb"])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            ctx.forward_device(dL, dab, dm, 0.5, want_dist=True, want_rgb=True, out_ab=out["ab"], out_dist=out["dist"], out_rgb=out["rgb"])
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("[%s] kernels only (graph replay x100, dist+rgb): %.1f us per forward, SM clock %d MHz"
              % (tag, e0.elapsed_time(e1) * 10.0, clk()))
        single = []
        for _ in range(30):
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            single.append(e0.elapsed_time(e1) * 1e3)
        print("[%s] kernels only, one replay at a time: p50 %.1f us" % (tag, pct(single, 50)))
    ctx.close()


def copies():
    h_in = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(720896, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(720896, dtype=torch.uint8, device="cuda")
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a, b, w = [], [], []
    for _ in range(30):
        t0 = time.perf_counter()
        e0.record(); d_in.copy_(h_in, non_blocking=True); e1.record(); h_out.copy_(d_out, non_blocking=True); e2.record()
        torch.cuda.synchronize()
        w.append((time.perf_counter() - t0) * 1e3)
        a.append(e0.elapsed_time(e1) * 1e3); b.append(e1.elapsed_time(e2) * 1e3)
    print("[copies] pinned H2D 1 MiB: p50 %.1f us   D2H 704 KiB: p50 %.1f us   both + sync, wall: p50 %.1f us"
          % (pct(a[5:], 50), pct(b[5:], 50), pct(w[5:], 50) * 1e3))


if __name__ == "__main__":
    copies()
    specs = sys.argv[1:] or ["default="]
    for spec in specs:
        name, _, body = spec.partition("=")
        opts = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in body.split(",") if kv}
        size = opts.pop("size", 256)               # size:512 = the high-resolution click
        run(name, opts, X=size)
