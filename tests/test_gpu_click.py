"""GPU: the click itself (BASELINE config 5) -- the pieces that ride on the click graph besides the conv trunk:
the resident image (idc_set_image: L uploaded once per photo), the announced click (idc_set_click: the clicked pixel's
pmf and K colour suggestions on the dist head's side branch, returned with the same graph launch) and the shared trunk
of the colour / distribution wrapper pair.  Everything must be bit-identical to the plain calls it short-cuts."""
import numpy as np
import pytest

from interactive_deep_colorization_b200 import _lib
from interactive_deep_colorization_b200 import colorize_image as CI
from oracle import synth
from tests import util

pytestmark = pytest.mark.gpu


def _clicks(n, X, seed):
    rs = np.random.RandomState(seed)
    return [(rs.randint(8, X - 8, 2), rs.uniform(-80, 80, 2)) for _ in range(n)]


def test_set_image_resident_l_matches_explicit_l(synth_sd):
    """forward_host(None, ...) after set_image == forward_host(L, ...), on the graph path (n = 1, pinned and pageable
    buffers) and on the chunked large-batch path (n = 8); without an image it is a state error, not a silent zero L."""
    X = 64
    L, ab, m = synth.synthetic_batch(8, X, seed=5, max_hints=4)
    ctx = util.make_ctx(synth_sd, X, X, max_n=8, dist=True)
    with pytest.raises(_lib.IdcError):
        ctx.forward_host(None, ab[:1], m[:1], 0.5)
    ref1 = ctx.forward_host(L[:1], ab[:1], m[:1], 0.5, want_dist=True, want_rgb=True, want_abq=True)
    ref8 = ctx.forward_host(L, ab, m, 0.5, want_dist=True, want_rgb=True)
    ctx.set_image(np.ascontiguousarray(L[:1]))
    got1 = ctx.forward_host(None, ab[:1], m[:1], 0.5, want_dist=True, want_rgb=True, want_abq=True)
    for k in ("ab", "dist", "rgb", "abq"):
        assert np.array_equal(ref1[k], got1[k]), k
    with pytest.raises(_lib.IdcError):          # one image resident, eight asked for
        ctx.forward_host(None, ab, m, 0.5)
    ctx.set_image(L)
    got8 = ctx.forward_host(None, ab, m, 0.5, want_dist=True, want_rgb=True)
    for k in ("ab", "dist", "rgb"):
        assert np.array_equal(ref8[k], got8[k]), k
    # pinned click buffers, hints only: [ab | mask] is one H2D
    ctx.set_image(np.ascontiguousarray(L[:1]))
    buf = ctx.click_buffers(1)
    buf["ab"][...] = ab[:1]; buf["mask"][...] = m[:1]
    ctx.set_dist_resident(True)
    for _ in range(2):                          # capture, then replay
        r = ctx.forward_host(None, buf["ab"], buf["mask"], 0.5, want_rgb=True, want_abq=True, out_ab=buf["out_ab"],
                             out_rgb=buf["out_rgb"], out_abq=buf["out_abq"])
        for k in ("ab", "rgb", "abq"):
            assert np.array_equal(ref1[k], r[k]), k
    assert np.array_equal(ctx.fetch_dist(0), ref1["dist"][0])
    ctx.set_image(None)
    with pytest.raises(_lib.IdcError):
        ctx.forward_host(None, buf["ab"], buf["mask"], 0.5, out_ab=buf["out_ab"])
    ctx.close()


def test_announced_click_returns_pmf_and_suggestions_with_the_forward(synth_sd):
    """idc_set_click: after the forward, fetch_dist / ab_reccs for the announced pixel are answered from pinned host
    memory -- bit-identical to the device-side calls on the same resident distribution; other pixels, another K or
    non-default clustering parameters fall back to the device; moving the click does not re-capture the graph."""
    X = 128
    L, ab, m = synth.synthetic_batch(1, X, seed=9, max_hints=0)
    ab = ab.copy(); m = m.copy()
    ctx = util.make_ctx(synth_sd, X, X, max_n=1, dist=True)
    ctx.set_dist_resident(True)
    buf = ctx.click_buffers(1)
    buf["L_mc"][...] = L
    kw = dict(want_rgb=True, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"])
    plain = util.make_ctx(synth_sd, X, X, max_n=1, dist=True)      # no click mode: every lookup goes to the device
    plain.set_dist_resident(True)
    for i, (loc, val) in enumerate(_clicks(6, X, 3)):
        CI.put_point(ab[0], m[0], loc, 3, val)
        buf["ab"][...] = ab; buf["mask"][...] = m
        y4, x4 = int(loc[0]) // 4, int(loc[1]) // 4
        K = (9, 5, 1)[i % 3]
        ctx.set_click(0, y4, x4, K)
        r = ctx.forward_host(buf["L_mc"], buf["ab"], buf["mask"], 0.5, **kw)
        p = plain.forward_host(L, ab, m, 0.5, want_rgb=True)
        assert np.array_equal(r["ab"], p["ab"]) and np.array_equal(r["rgb"], p["rgb"])
        want_pmf = plain.fetch_dist(0, y4, x4)
        assert np.array_equal(ctx.fetch_dist(0, y4, x4), want_pmf)
        assert abs(float(want_pmf.sum()) - 1.0) < 1e-4
        cw, fw, iw = plain.ab_reccs(0, y4, x4, K=K)
        cg, fg, ig = ctx.ab_reccs(0, y4, x4, K=K)
        assert np.array_equal(cg, cw) and np.array_equal(fg, fw) and ig == iw
        # explicit default grid == NULL grid (the wrapper always passes pts_in_hull)
        g = np.arange(-110, 120, 10)
        pts = np.array(np.meshgrid(g, g)).reshape((2, 529)).T
        cg2, fg2, _ = ctx.ab_reccs(0, y4, x4, K=K, pts=pts)
        assert np.array_equal(cg2, cw) and np.array_equal(fg2, fw)
        # fall-backs: another pixel, another K, non-default restarts
        y2, x2 = (y4 + 3) % (X // 4), (x4 + 5) % (X // 4)
        assert np.array_equal(ctx.fetch_dist(0, y2, x2), plain.fetch_dist(0, y2, x2))
        for kwargs in (dict(K=K + 1), dict(K=K, n_init=4), dict(K=K, max_iter=3)):
            a, b = ctx.ab_reccs(0, y4, x4, **kwargs), plain.ab_reccs(0, y4, x4, **kwargs)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    # mode off again: same numbers through the device path
    ctx.set_click(0, -1, 0, 0)
    r = ctx.forward_host(buf["L_mc"], buf["ab"], buf["mask"], 0.5, **kw)
    assert np.array_equal(r["ab"], p["ab"])
    assert np.array_equal(ctx.fetch_dist(0, y4, x4), want_pmf)
    ctx.close(); plain.close()


def test_announced_click_in_a_small_batch(synth_sd):
    """The click may sit in any image of a graph-path batch (n <= 4); an image index outside the batch is answered by
    the device path's own error, not by stale host data."""
    X = 64
    L, ab, m = synth.synthetic_batch(3, X, seed=21, max_hints=5)
    ctx = util.make_ctx(synth_sd, X, X, max_n=3, dist=True)
    ctx.set_dist_resident(True)
    plain = util.make_ctx(synth_sd, X, X, max_n=3, dist=True)
    plain.set_dist_resident(True)
    plain.forward_host(L, ab, m, 0.5)
    for img, y4, x4, K in ((0, 3, 5, 4), (2, 15, 0, 9), (1, 7, 7, 0)):
        ctx.set_click(img, y4, x4, K)
        r = ctx.forward_host(L, ab, m, 0.5)
        assert np.array_equal(r["ab"], plain.forward_host(L, ab, m, 0.5)["ab"])
        assert np.array_equal(ctx.fetch_dist(img, y4, x4), plain.fetch_dist(img, y4, x4))
        if K:
            a, b = ctx.ab_reccs(img, y4, x4, K=K), plain.ab_reccs(img, y4, x4, K=K)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    ctx.set_click(3, 1, 1, 5)                    # image 3 of a 3-image batch does not exist
    ctx.forward_host(L, ab, m, 0.5)
    with pytest.raises(_lib.IdcError):
        ctx.fetch_dist(3, 1, 1)
    ctx.close(); plain.close()


def test_click_at_512_against_the_oracle(synth_sd):
    """The interactive plan at the high-resolution size of BASELINE config 4 (512^2, batch 1: 64^2 bottleneck, other
    split-K shapes than at 256^2): raw ab within tolerance of the oracle, dist normalised, announced click identical to
    the plain lookups, graph replay identical."""
    X = 512
    L, ab, m = synth.synthetic_batch(1, X, seed=33, max_hints=8)
    ctx = util.make_ctx(synth_sd, X, X, max_n=1, dist=True)
    ctx.set_dist_resident(True)
    reg, dist = util.oracle_forward(synth_sd, L, ab, m, 0.5, dist=True)
    ctx.set_click(0, 70, 101, 9)
    r = ctx.forward_host(L, ab, m, 0.5, want_rgb=True)
    r2 = ctx.forward_host(L, ab, m, 0.5, want_rgb=True)
    assert np.array_equal(r["ab"], r2["ab"]) and np.array_equal(r["rgb"], r2["rgb"])
    err = util.maxabs(r["ab"], reg)
    print("512^2 click: max|d ab| vs oracle = %.3e" % err)
    assert err <= 1e-3
    pmf = ctx.fetch_dist(0, 70, 101)
    assert util.maxabs(pmf, dist.numpy()[0, :, 70, 101]) < 1e-5 and abs(float(pmf.sum()) - 1.0) < 1e-4
    full = ctx.fetch_dist(0)
    assert np.array_equal(full[:, 70, 101], pmf)
    served = ctx.ab_reccs(0, 70, 101, K=9)
    ctx.set_click(0, -1, 0, 0)
    ctx.forward_host(L, ab, m, 0.5, want_rgb=True)
    plain = ctx.ab_reccs(0, 70, 101, K=9)
    assert np.array_equal(served[0], plain[0]) and np.array_equal(served[1], plain[1])
    ctx.close()


def test_shared_trunk_pair_is_one_forward_per_click(synth_sd):
    """launcher --backend b200: colour model and distribution model share one context (ideepcolor.py:34-38 loads the
    same checkpoint into both).  Per click the pair must publish exactly what two separately prepared models publish,
    with ONE forward -- whichever of the two the GUI calls first."""
    X = 64
    img = np.random.RandomState(2).randint(0, 256, (X, X, 3)).astype(np.uint8)

    def pair(shared):
        cm = CI.ColorizeImageB200(Xd=X, maskcent=True)
        cm.prep_net(state_dict=synth_sd, dist=shared)
        cd = CI.ColorizeImageB200Dist(Xd=X, maskcent=True)
        if shared:
            cd.share_trunk(cm)
        else:
            cd.prep_net(state_dict=synth_sd)
        cm.set_image(img); cd.set_image(img.copy())
        return cm, cd
    (cm_s, cd_s), (cm_p, cd_p) = pair(True), pair(False)
    ctx = cm_s.net._context(X, X, 1)
    calls = []
    inner = ctx.forward_host
    ctx.forward_host = lambda *a, **kw: (calls.append(1), inner(*a, **kw))[1]
    ab64, m64 = np.zeros((2, X, X)), np.zeros((1, X, X))
    for i, (loc, val) in enumerate(_clicks(6, X, 7)):
        CI.put_point(ab64, m64, loc, 2, val)
        h, w = int(loc[0]), int(loc[1])
        n0 = len(calls)
        if i % 2 == 0:                       # compute_result then predict_color (ui/gui_draw.py:272, :250)
            cd_s.hint_click(h, w, K=5)
            rgb_s = cm_s.net_forward(ab64, m64)
            ret_s = cd_s.net_forward(ab64.copy(), m64.copy())
        else:                                # the per-click hook calls predict_color first (launcher.py)
            cd_s.hint_click(None, None)
            ret_s = cd_s.net_forward(ab64, m64)
            rgb_s = cm_s.net_forward(ab64.copy(), m64.copy())
        assert len(calls) - n0 == 1, "the shared pair ran %d forwards for one click" % (len(calls) - n0)
        rgb_p = cm_p.net_forward(ab64, m64)
        ret_p = cd_p.net_forward(ab64, m64)
        assert np.array_equal(rgb_s, rgb_p) and np.array_equal(cm_s.output_ab, cm_p.output_ab)
        assert np.array_equal(cm_s.output_ab_raw, cm_p.output_ab_raw) and np.array_equal(ret_s, ret_p)
        assert np.array_equal(np.asarray(cd_s.dist_ab[:, h, w]), np.asarray(cd_p.dist_ab[:, h, w]))
        a_s, c_s = cd_s.get_ab_reccs(h, w, K=5, return_conf=True)
        a_p, c_p = cd_p.get_ab_reccs(h, w, K=5, return_conf=True)
        assert np.array_equal(a_s, a_p) and np.array_equal(c_s, c_p)
    # a different image on one side must not be answered from the other side's forward
    cd_s.set_image(img[::-1].copy())
    n0 = len(calls)
    ret2 = cd_s.net_forward(ab64, m64)
    assert len(calls) - n0 == 1 and not np.array_equal(ret2, ret_s)
    cd_p.set_image(img[::-1].copy())
    assert np.array_equal(ret2, cd_p.net_forward(ab64, m64))
    rgb3 = cm_s.net_forward(ab64, m64)          # colour model still holds the first image
    assert len(calls) - n0 == 2 and np.array_equal(rgb3, cm_p.net_forward(ab64, m64))
