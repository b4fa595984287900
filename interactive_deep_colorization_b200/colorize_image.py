"""Host-side mirror of the reference wrapper classes, backed by the B200 engine.

API kept verbatim from /root/reference/data/colorize_image.py so `ideepcolor.py:62-72`, the Qt
GUI (`ui/gui_draw.py:109-113,258-286`) and the notebooks work unchanged:

    ColorizeImageBase      (:39-198)   image prep, getters, full-res zoom
    ColorizeImageB200      <-> ColorizeImageTorch      (:201-276)
    ColorizeImageB200Dist  <-> ColorizeImageTorchDist  (:279-372)
    ColorizeImageB200GlobDist <-> ColorizeImageCaffeGlobDist (:445-463) semantics on the torch scaling

Differences, all deliberate: no matplotlib / scikit-image imports (colour math in .color);
`prep_net` also accepts an in-memory `state_dict`; the network forward, the 529-bin softmax
and the Lab->RGB post-process run in libidc_b200.so.  There is no CPU fallback.
"""
import numpy as np

from . import color
from .color import lab2rgb_transpose, rgb2lab_transpose  # noqa: F401  (re-exported like the reference)


def put_point(input_ab, mask, loc, p, val):
    """Notebook helper (DemoInteractiveColorization.ipynb:131-139): paint a (2p+1)^2 hint."""
    input_ab[:, loc[0] - p:loc[0] + p + 1, loc[1] - p:loc[1] + p + 1] = np.array(val)[:, np.newaxis, np.newaxis]
    mask[:, loc[0] - p:loc[0] + p + 1, loc[1] - p:loc[1] + p + 1] = 1
    return (input_ab, mask)


def _zoom(a, factors, order):
    from scipy.ndimage import zoom
    return zoom(a, factors, order=order)


class ColorizeImageBase(object):
    """Image state + getters.  Attribute and method names are the reference's public surface
    (data/colorize_image.py:39-198); the bodies are organised around three helpers:
    `_ingest` (RGB -> Lab planes), `_to_fullres` (scipy zoom to the full-resolution grid) and
    `_render` (Lab planes -> uint8 RGB)."""

    def __init__(self, Xd=256, Xfullres_max=10000):
        self.Xd, self.Xfullres_max = Xd, Xfullres_max
        self.img_l_set = self.net_set = self.img_just_set = False

    def prep_net(self):
        raise Exception("Should be implemented by base class")

    # ----- image prep: reference load_image :52-66, set_image :68-77 -----
    def _ingest(self, rgb_full, rgb_net):
        self.img_rgb_fullres = rgb_full
        self._set_img_lab_fullres_()
        self.img_rgb = rgb_net
        self._set_img_lab_()
        self._set_img_lab_mc_()

    def load_image(self, input_path):
        import cv2
        bgr = cv2.imread(input_path, 1)
        if bgr is None:
            raise IOError("cannot read image %r" % (input_path,))
        full = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        self._ingest(full.copy(), cv2.resize(full, (self.Xd, self.Xd)).copy())   # cv2 default = bilinear

    def set_image(self, input_image):
        self._ingest(input_image.copy(), input_image)

    # ----- forward preconditions + hint normalisation: reference :79-96 -----
    def net_forward(self, input_ab, input_mask):
        for ok, what in ((self.img_l_set, 'an image'), (self.net_set, 'a net')):
            if not ok:
                print('I need to have %s!' % what)
                return -1
        self.input_ab, self.input_mask = input_ab, input_mask
        # reference :92-93.  With the PyTorch constants (ab_mean 0, ab_norm 1, mask_mult 1) both statements are exact
        # identities; skipping the two float64 temporaries (2.5 MB of numpy traffic) is worth ~0.1 ms per click
        self.input_ab_mc = input_ab if (self.ab_mean == 0 and self.ab_norm == 1) else (input_ab - self.ab_mean) / self.ab_norm
        self.input_mask_mult = input_mask if self.mask_mult == 1 else input_mask * self.mask_mult
        return 0

    def get_result_PSNR(self, result=-1, return_SE_map=False):
        use_own = np.array((result)).flatten()[0] == -1
        err2 = (1. * self.img_rgb - (self.get_img_forward() if use_own else result.copy())) ** 2
        psnr = 20 * np.log10(255. / np.sqrt(np.mean(err2)))
        return (psnr, err2) if return_SE_map else psnr

    # ----- rendering helpers -----
    @staticmethod
    def _render(l_plane, ab_planes=None):
        if ab_planes is None:
            ab_planes = np.zeros((2,) + tuple(l_plane.shape[1:]))
        return lab2rgb_transpose(l_plane, ab_planes)

    def _to_fullres(self, planes, like, order):
        fh = 1. * self.img_l_fullres.shape[1] / like.shape[1]
        fw = 1. * self.img_l_fullres.shape[2] / like.shape[2]
        return _zoom(planes, (1, fh, fw), order)

    # ----- getters: reference :111-158 -----
    def get_img_forward(self):
        return self.output_rgb

    def get_img_gray(self):
        return self._render(self.img_l)

    def get_img_gray_fullres(self):
        return self._render(self.img_l_fullres)

    def get_img_fullres(self):       # bilinear up-zoom of the (quantised) output ab, then Lab->RGB
        return self._render(self.img_l_fullres, self._to_fullres(self.output_ab, self.output_ab, 1))

    def get_input_img_fullres(self):
        return self._render(self.img_l_fullres, self._to_fullres(self.input_ab, self.input_ab, 1))

    def get_input_img(self):
        return self._render(self.img_l, self.input_ab)

    def get_img_mask(self):
        return self._render(100. * (1 - self.input_mask))

    def get_img_mask_fullres(self):
        return self._render(100. * (1 - self._to_fullres(self.input_mask, self.input_ab, 0)))

    def get_sup_img(self):
        return self._render(50 * self.input_mask, self.input_ab)

    def get_sup_fullres(self):
        return self._render(50 * self._to_fullres(self.input_mask, self.output_ab, 0),
                            self._to_fullres(self.input_ab, self.output_ab, 0))

    # ----- Lab planes: reference :161-198 -----
    @staticmethod
    def _lab_planes(rgb):
        lab = color.rgb2lab(rgb).transpose((2, 0, 1))
        return lab, lab[[0]], lab[1:]

    def _set_img_lab_fullres_(self):
        big = max(self.img_rgb_fullres.shape[:2])
        if big > self.Xfullres_max:          # cap the longest side
            z = 1. * self.Xfullres_max / big
            self.img_rgb_fullres = _zoom(self.img_rgb_fullres, (z, z, 1), 1)
        self.img_lab_fullres, self.img_l_fullres, self.img_ab_fullres = self._lab_planes(self.img_rgb_fullres)

    def _set_img_lab_(self):
        self.img_lab, self.img_l, self.img_ab = self._lab_planes(self.img_rgb)

    def _set_img_lab_mc_(self):
        div = np.array((self.l_norm, self.ab_norm, self.ab_norm), dtype=np.float64).reshape(3, 1, 1)
        sub = np.array((self.l_mean, self.ab_mean, self.ab_mean), dtype=np.float64).reshape(3, 1, 1) / div
        self.img_lab_mc = self.img_lab / div - sub
        self._set_img_l_()

    def _set_img_l_(self):
        self.img_l_mc = self.img_lab_mc[[0]]
        self.img_l_set = True

    def _set_img_ab_(self):
        self.img_ab_mc = self.img_lab_mc[[1, 2]]

    def _set_out_ab_(self):
        # output_ab is re-derived from the uint8 RGB, i.e. quantised (reference :196-198)
        self.output_lab = rgb2lab_transpose(self.output_rgb)
        self.output_ab = self.output_lab[1:]


class ColorizeImageB200(ColorizeImageBase):
    """<-> ColorizeImageTorch (reference :201-276)."""

    def __init__(self, Xd=256, maskcent=False, engine="tcgen05", fast_fp16=False, gpu_prepost=True):
        print('ColorizeImageB200 instantiated')
        self.gpu_prepost = gpu_prepost    # quantised output_ab and the full-res render on the GPU (row f1)
        ColorizeImageBase.__init__(self, Xd)
        self.l_norm = 1.
        self.ab_norm = 1.
        self.l_mean = 50.
        self.ab_mean = 0.
        self.mask_mult = 1.
        self.mask_cent = .5 if maskcent else 0
        self.engine = engine
        self.fast_fp16 = fast_fp16
        # torch-path bin grid (reference :213; (b,a)-ordered meshgrid)
        self.pts_in_hull = np.array(np.meshgrid(np.arange(-110, 120, 10), np.arange(-110, 120, 10))).reshape((2, 529)).T

    def prep_net(self, gpu_id=None, path='', dist=False, state_dict=None):
        import torch
        from .model import SIGGRAPHGeneratorB200
        print('path = %s' % path)
        print('Model set! dist mode? ', dist)
        self.net = SIGGRAPHGeneratorB200(dist=dist, device=0 if gpu_id is None else int(gpu_id), engine=self.engine,
                                         fast_fp16=self.fast_fp16)
        if state_dict is None:
            state_dict = torch.load(path, map_location='cpu')
        if hasattr(state_dict, '_metadata'):
            del state_dict._metadata
        self.net.load_state_dict(state_dict)
        self.net.cuda()
        self.net.eval()
        self.net_set = True

    def net_forward(self, input_ab, input_mask):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        ctx = self.net._context(self.img_l_mc.shape[-2], self.img_l_mc.shape[-1], 1)
        # ONE C-ABI call and one round trip: H2D, forward, fused Lab->RGB post-process (reference :263-264) and the
        # quantised output_ab = rgb2lab(output_rgb)[1:] (reference :267 -> :196-198) in the same kernel, D2H
        self._click(ctx, float(self.mask_cent))
        return self.output_rgb

    def _click(self, ctx, maskcent, glob=None, mask_div=1.0, want_rgb=True, publish_rgb=None, need_dist=False):
        """Stage the reference's float64 arrays into the context's page-locked click buffers (the float64 -> float32
        conversion IS the only CPU copy; L is re-staged only when the image changed), run idc_forward_host_q with
        the pinned buffers (zero-copy graph path) and publish copies of the results as the reference's attributes."""
        buf = self._click_buffers(ctx, glob is not None)
        want_q = bool(want_rgb and self.gpu_prepost)
        if ctx._wrapper_shared and glob is None and \
                self._same_as_last_forward(ctx, buf, maskcent, mask_div, want_rgb, want_q, need_dist):
            # share_trunk: the other model of the pair just ran this very forward; its results are still in the buffers
            r = {"ab": buf["out_ab"], "rgb": buf["out_rgb"], "abq": buf["out_abq"]}
        else:
            self._stage_image(ctx, buf)
            np.copyto(buf["ab"][0], self.input_ab_mc, casting='unsafe')
            if mask_div == 1.0:
                np.copyto(buf["mask"][0], self.input_mask_mult, casting='unsafe')
            else:
                np.divide(self.input_mask_mult, mask_div, out=buf["mask"][0], casting='unsafe')
            if glob is not None:
                buf["glob"][...] = glob
            ctx._wrapper_last = None
            # L_mc = None: the image uploaded by _stage_image (idc_set_image) -- a click moves only the hints
            r = ctx.forward_host(None, buf["ab"], buf["mask"], maskcent, glob=buf["glob"], want_rgb=want_rgb,
                                 want_abq=want_q, out_ab=buf["out_ab"], out_rgb=buf["out_rgb"] if want_rgb else None,
                                 out_abq=buf["out_abq"] if want_q else None)
            ctx._wrapper_last = (float(maskcent), float(mask_div), glob is not None, bool(want_rgb), want_q,
                                 bool(getattr(ctx, "_dist_resident", False)))
        self.output_ab_raw = r["ab"][0].copy()   # raw net output (the parity quantity)
        if want_rgb and (publish_rgb is None or publish_rgb):
            self.output_rgb = r["rgb"][0].copy()
            if want_q:
                self.output_ab = r["abq"][0].copy()
                self._output_lab = None          # output_lab (L plane included) is derived on demand
            else:
                ColorizeImageBase._set_out_ab_(self)
        return r

    @staticmethod
    def _click_buffers(ctx, with_glob):
        buf = getattr(ctx, "_wrapper_click", None)
        if buf is None or with_glob != (buf["glob"] is not None):
            buf = ctx.click_buffers(1, glob=with_glob)
            ctx._wrapper_click = buf
            ctx._wrapper_staged_l = []
            ctx._wrapper_last = None
        return buf

    def _same_as_last_forward(self, ctx, buf, maskcent, mask_div, want_rgb, want_q, need_dist=False):
        """Did the (shared) context just run exactly this image + these hints (float32, as staged), producing at least
        the outputs asked for (RGB, quantised ab, the resident distribution)?"""
        last = ctx._wrapper_last
        if last is None or last[:3] != (float(maskcent), float(mask_div), False) or (want_rgb and not last[3]) or \
                (want_q and not last[4]) or (need_dist and not last[5]):
            return False
        if not any(a is self.img_l_mc for a in ctx._wrapper_staged_l):
            l32 = np.ascontiguousarray(self.img_l_mc, dtype=np.float32).reshape(buf["L_mc"].shape)
            if not (ctx._wrapper_staged_l and np.array_equal(buf["L_mc"], l32)):
                return False
            ctx._wrapper_staged_l = ctx._wrapper_staged_l[-3:] + [self.img_l_mc]
        mask32 = np.asarray(self.input_mask_mult, dtype=np.float32)
        if mask_div != 1.0:
            mask32 = mask32 / np.float32(mask_div)
        return (np.array_equal(buf["ab"][0], np.asarray(self.input_ab_mc, dtype=np.float32)) and
                np.array_equal(buf["mask"][0], mask32))

    def _stage_image(self, ctx, buf):
        """The L plane goes to the device once per image (reference: set_image / load_image, :68-77), not once per
        click.  `_wrapper_staged_l` lists the img_l_mc arrays known to equal the resident plane (several wrapper
        objects may share one context, see ColorizeImageB200Dist.share_trunk)."""
        if any(a is self.img_l_mc for a in ctx._wrapper_staged_l):
            return
        l32 = np.ascontiguousarray(self.img_l_mc, dtype=np.float32).reshape(buf["L_mc"].shape)
        if ctx._wrapper_staged_l and np.array_equal(buf["L_mc"], l32):
            ctx._wrapper_staged_l = ctx._wrapper_staged_l[-3:] + [self.img_l_mc]
            return
        buf["L_mc"][...] = l32
        ctx.set_image(buf["L_mc"])
        ctx._wrapper_staged_l = [self.img_l_mc]
        ctx._wrapper_last = None

    @property
    def output_lab(self):
        """reference :197 (`self.output_lab = rgb2lab_transpose(self.output_rgb)`); nothing in the reference reads it
        besides `_set_out_ab_` itself, so the fused path computes it lazily."""
        if getattr(self, "_output_lab", None) is None:
            self._output_lab = rgb2lab_transpose(self.output_rgb)
        return self._output_lab

    @output_lab.setter
    def output_lab(self, v):
        self._output_lab = v

    def get_img_forward(self):
        return self.output_rgb

    def get_img_gray(self):
        return lab2rgb_transpose(self.img_l, np.zeros((2, self.Xd, self.Xd)))

    # ----- row f1, image-load side: reference load_image :52-66 on the GPU when a net is set -----
    def _ingest(self, rgb_full, rgb_net):
        """Full-resolution rgb2lab (the reference spends seconds of float64 numpy on an 18 MP photo, :161-170) stays in
        HBM as a DeviceLab; only the Xd x Xd planes come back.  `rgb_net` None = resize here with the cv2-exact kernel."""
        big = max(rgb_full.shape[:2])
        if not (self.gpu_prepost and self.net_set) or big > self.Xfullres_max or rgb_full.dtype != np.uint8:
            if rgb_net is None:
                import cv2
                rgb_net = cv2.resize(rgb_full, (self.Xd, self.Xd)).copy()
            return ColorizeImageBase._ingest(self, rgb_full, rgb_net)
        from . import prepost
        small, lab, dlab = prepost.load_image_gpu(rgb_full, self.Xd, self._device())
        self.img_rgb_fullres = rgb_full
        self.img_lab_fullres, self.img_l_fullres, self.img_ab_fullres = dlab, dlab.view(slice(0, 1)), dlab.view(slice(1, 3))
        if rgb_net is None:
            self.img_rgb = small
            self.img_lab = lab
        else:                                             # set_image: the caller pre-resized (reference :68-77)
            self.img_rgb = rgb_net
            self.img_lab = prepost.rgb2lab_gpu(rgb_net, self._device()) if rgb_net.dtype == np.uint8 else self._lab_planes(rgb_net)[0]
        self.img_l, self.img_ab = self.img_lab[[0]], self.img_lab[1:]
        self._set_img_lab_mc_()

    def load_image(self, input_path):
        import cv2
        bgr = cv2.imread(input_path, 1)
        if bgr is None:
            raise IOError("cannot read image %r" % (input_path,))
        self._ingest(np.ascontiguousarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)), None)

    # ----- row f1: the numpy/scipy steps either side of the network, on the GPU when a net is set -----
    def _device(self):
        """CUDA device ordinal of the engine behind this wrapper (the Torch-named classes hold a module in
        `self.net`, the GlobDist / Caffe-named ones an LhnContext in `self._ctx`)."""
        ctx = getattr(self, "_ctx", None)
        return ctx.device if ctx is not None else self.net.b200_device

    def _set_out_ab_(self):
        if not (self.gpu_prepost and self.net_set):
            return ColorizeImageBase._set_out_ab_(self)
        from . import prepost
        self.output_lab = prepost.rgb2lab_gpu(self.output_rgb, self._device())
        self.output_ab = self.output_lab[1:]

    def get_img_fullres(self):
        if not (self.gpu_prepost and self.net_set):
            return ColorizeImageBase.get_img_fullres(self)
        from . import prepost
        return prepost.fullres_rgb_gpu(self.output_ab, self.img_l_fullres, self._device())


class _LazyUpsampledDist(object):
    """[529, X, X] view of the [529, X/4, X/4] distribution.  The reference materialises the nearest x4
    upsample (139 MB at 256^2, model.py:160) and copies it to the host; its consumers only index single
    pixels (`dist_ab[:, h, w]`, data/colorize_image.py:329).  Here the distribution stays on the device
    (`fetch` pulls 529 floats per lookup) or is a host [529, X/4, X/4] array; either way it is
    replicated on read."""

    def __init__(self, d64=None, fetch=None, shape64=None):
        self.d64, self._fetch = d64, fetch
        s = d64.shape if d64 is not None else shape64
        self.shape = (s[0], s[1] * 4, s[2] * 4)
        self.dtype = np.dtype(np.float32)

    def _plane(self):
        if self.d64 is None:
            self.d64 = self._fetch(None, None)
        return self.d64

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 3 and all(isinstance(i, (int, np.integer)) for i in idx[1:]):
            if self.d64 is None:
                return self._fetch(int(idx[1]) // 4, int(idx[2]) // 4)[idx[0]]
            return self.d64[idx[0], idx[1] // 4, idx[2] // 4]
        return self.__array__()[idx]

    def __array__(self, dtype=None, copy=None):
        a = np.repeat(np.repeat(self._plane(), 4, axis=1), 4, axis=2)
        return a.astype(dtype) if dtype is not None else a


class ColorizeImageB200Dist(ColorizeImageB200):
    """<-> ColorizeImageTorchDist (reference :279-372)."""

    def __init__(self, Xd=256, maskcent=False, engine="tcgen05", fast_fp16=False, materialize_full=False):
        ColorizeImageB200.__init__(self, Xd, engine=engine, fast_fp16=fast_fp16)
        self.dist_ab_set = False
        self.pts_grid = np.array(np.meshgrid(np.arange(-110, 120, 10), np.arange(-110, 120, 10))).reshape((2, 529)).T
        self.in_hull = np.ones(529, dtype=bool)
        self.AB = self.pts_grid.shape[0]
        self.A = int(np.sqrt(self.AB))
        self.B = int(np.sqrt(self.AB))
        self.materialize_full = materialize_full
        self.dist_entropy = np.zeros((self.Xd, self.Xd))
        self.mask_cent = .5 if maskcent else 0

    def prep_net(self, gpu_id=None, path='', dist=True, S=.2, state_dict=None):
        ColorizeImageB200.prep_net(self, gpu_id=gpu_id, path=path, dist=dist, state_dict=state_dict)

    def share_trunk(self, color_model):
        """ONE forward per click instead of two.  The PyTorch backend loads the same checkpoint into both models
        (ideepcolor.py:34-38 "same model used for both") and the GUI feeds both the same hints, one after the other
        (ui/gui_draw.py:250-258 predict_color, :272-279 compute_result); the distribution head hangs off conv8_3 of
        the very trunk the colour model just ran.  After `share_trunk(color_model)` -- color_model prepared with
        `prep_net(..., dist=True)` -- this object uses the colour model's network: when net_forward sees the image and
        hints the shared context ran last, it publishes that forward's outputs (the resident distribution, the raw ab
        map) without launching anything; otherwise it runs the forward itself -- and the colour model's next
        net_forward with the same hints is answered from THAT forward (the sharing is symmetric, whichever model the
        GUI calls first pays).  Replaces prep_net."""
        net = getattr(color_model, "net", None)
        if not getattr(color_model, "net_set", False) or net is None or not getattr(net, "dist", False):
            raise ValueError("share_trunk: prepare the colour model with prep_net(..., dist=True) first")
        self.net = net
        self.net_set = True
        self._trunk = color_model
        Xd = self.Xd
        ctx = net._context(Xd, Xd, 1)
        ctx.set_dist_resident(True)          # every forward of the shared context keeps its distribution
        ctx._wrapper_shared = True           # _click may now answer from the other model's forward
        return self

    def hint_click(self, h, w, K=9):
        """Announce the pixel the GUI is about to ask suggestions for (ui/gui_draw.py:184 `suggest_color(h=y, w=x, K=9)`)
        BEFORE the forward: its pmf and the K suggestions then come back with the click itself (idc_set_click), and
        `dist_ab[:, h, w]` / `get_ab_reccs(h, w, K)` cost no device work.  h = None switches it off."""
        ctx = self.net._context(self.Xd, self.Xd, 1)
        if h is None:
            ctx.set_click(0, -1, 0, 0)
        else:
            ctx.set_click(0, int(h) // 4, int(w) // 4, int(K))

    def net_forward(self, input_ab, input_mask):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        Xh, Xw = self.img_l_mc.shape[-2], self.img_l_mc.shape[-1]
        ctx = self.net._context(Xh, Xw, 1)
        if self.materialize_full:
            A = np.ascontiguousarray(self.img_l_mc, dtype=np.float32)[None]
            B = np.ascontiguousarray(self.input_ab_mc, dtype=np.float32)[None]
            M = np.ascontiguousarray(self.input_mask_mult, dtype=np.float32)[None]
            r = ctx.forward_host(A, B, M, float(self.mask_cent), want_dist=True)
            self.dist_ab_64 = r["dist"][0]                   # [529, X/4, X/4]
            self.output_ab_raw = r["ab"][0]
        else:
            ctx.set_dist_resident(True)                      # dist stays in HBM; pixels are fetched on demand
            # on a shared context keep the colour model's graph (same outputs requested -> no re-capture); the
            # distribution model itself publishes no RGB (reference :297-320 never sets output_rgb)
            self._click(ctx, float(self.mask_cent), want_rgb=getattr(self, "_trunk", None) is not None, publish_rgb=False,
                        need_dist=True)
        if self.materialize_full:
            self.dist_ab = np.repeat(np.repeat(self.dist_ab_64, 4, axis=1), 4, axis=2)
            self.dist_ab_full = np.zeros((self.AB, self.Xd, self.Xd))
            self.dist_ab_full[self.in_hull, :, :] = self.dist_ab
            self.dist_ab_grid = self.dist_ab_full.reshape((self.A, self.B, self.Xd, self.Xd))
        else:
            self.dist_ab = _LazyUpsampledDist(fetch=lambda y4, x4: ctx.fetch_dist(0, y4, x4),
                                              shape64=(529, Xh // 4, Xw // 4))
        self._dist_ctx = ctx
        self.dist_ab_set = True
        # reference returns the regression output scaled by 110 twice (model.py:166-168, q1)
        return self.output_ab_raw * 110.0

    def get_ab_reccs(self, h, w, K=5, N=25000, return_conf=False, method='gpu'):
        """Colour suggestions at pixel (h, w) (reference :322-354).  The reference draws N samples from the
        529-bin distribution by inverse-CDF lookup, k-means them and orders the clusters by occupancy.
        method='gpu' (default): the N -> infinity limit of that, weighted k-means on the device over the
        resident distribution (idc_ab_reccs; deterministic, N unused).  method='sampled': the reference's
        stochastic procedure on the host (np.random + sklearn), for side-by-side comparison."""
        if not self.dist_ab_set:
            print('Need to set prediction first')
            return 0
        if method == 'gpu':
            if not self.materialize_full:                    # the pixel's pmf never leaves the device
                centers, conf, _ = self._dist_ctx.ab_reccs(0, int(h) // 4, int(w) // 4, K=K, pts=self.pts_in_hull)
            else:
                from .prepost import ab_reccs_pmf_gpu
                centers, conf, _ = ab_reccs_pmf_gpu(np.asarray(self.dist_ab[:, h, w]), K=K, pts=self.pts_in_hull,
                                                    device=self._dist_ctx.device)
            centers, conf = centers.astype(np.float64), conf.astype(np.float64)
            return (centers, conf) if return_conf else centers
        if method != 'sampled':
            raise ValueError("method must be 'gpu' or 'sampled'")
        from sklearn.cluster import KMeans
        cdf = np.cumsum(np.asarray(self.dist_ab[:, h, w]))
        cdf /= cdf[-1]
        u = np.random.uniform(low=0, high=1.0, size=N)
        samples = self.pts_in_hull[np.searchsorted(cdf, u, side='right'), :]   # == np.digitize(u, cdf)
        km = KMeans(n_clusters=K).fit(samples)
        mass = np.bincount(km.labels_, minlength=K)
        order = np.argsort(mass)[::-1]
        centers, conf = km.cluster_centers_[order, :], mass[order] / float(N)
        return (centers, conf) if return_conf else centers

    def compute_entropy(self):
        d = np.asarray(self.dist_ab)
        self.dist_entropy = np.sum(d * np.log(d), axis=0)


class ColorizeImageB200GlobDist(ColorizeImageB200):
    """<-> ColorizeImageCaffeGlobDist (reference :445-463): colorization conditioned on a global ab histogram.
    The global-hints branch of models/global_model/deploy_nodist.prototxt:38-172,501-527 is bolted onto the
    local-hints network (a superset of the Caffe global model, whose conv1_1 ignores the local hints); its
    weights arrive as extra state_dict keys `glob.{0..3}.*` (include/idc_b200.h).  Spec-only: no reference
    weights or vectors exist for it offline."""

    def __init__(self, Xd=256, maskcent=False, engine="tcgen05"):
        ColorizeImageB200.__init__(self, Xd, maskcent=maskcent, engine=engine)
        self.glob_mask_mult = 1.

    def prep_net(self, gpu_id=None, path='', state_dict=None):
        import torch
        from .engine import LhnContext
        if state_dict is None:
            state_dict = torch.load(path, map_location='cpu')
        self._ctx = LhnContext(device=0 if gpu_id is None else int(gpu_id), max_n=1, H=self.Xd, W=self.Xd,
                               engine=self.engine, global_hints=True)
        self._ctx.load_state_dict(state_dict)
        self.net_set = True

    def get_global_histogram(self, ref_rgb_u8):
        """DemoGlobalHistogramTransfer.ipynb:176-182: the reference image is resized to Xd x Xd, then
        global_stats.prototxt -> the first 313 entries are `glob_dist`."""
        import cv2
        from . import prepost
        img = cv2.resize(ref_rgb_u8, (self.Xd, self.Xd))
        self.glob_vec = prepost.global_stats_gpu(img, self._ctx.device)
        return self.glob_vec[:313].copy()

    def net_forward(self, input_ab, input_mask, glob_dist=-1):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        glob = np.zeros((1, 316), np.float32)            # "run without this, zero it out" (reference :454-456)
        if np.array(glob_dist).flatten()[0] != -1:
            glob[0, :313] = np.asarray(glob_dist, dtype=np.float32)
            glob[0, 313] = self.glob_mask_mult             # reference :458-459; the s_avg input stays 0 as in the reference
        self._click(self._ctx, float(self.mask_cent), glob=glob)
        return self.output_rgb


# =============================================================================================
# Caffe-named wrapper surface (reference :375-442, :445-463, :466-561).  The notebooks and ideepcolor.py:60-65
# instantiate ColorizeImageCaffe / ColorizeImageCaffeDist / ColorizeImageCaffeGlobDist; these classes keep those
# names' semantics on the B200 engine:
#   * Caffe scaling: the deploy nets feed RAW L-50, raw ab and mask x 110 into conv1_1 and scale the
#     regression head by 100 (deploy_nodist.prototxt:19-51, :812-822; `self.mask_mult = 110.`, :383), where the
#     PyTorch model feeds L/100, ab/110, mask and scales by 110.  The engine normalises the PyTorch way inside
#     conv1_1_kernel, so a Caffe-scaled weight set is mapped exactly at load time (conv1_1 input channels x 100, x 110,
#     x 110; `tanh_scale` = 100) -- see `caffe_scaled_state_dict`.
#   * 313-bin head (deploy_nopred.prototxt:651-850): `dist_ab` = softmax(S * logits) over the in-gamut bins,
#     `pred_ab` = annealed mean (T = 2.6); `get_ab_reccs` works on `pts_in_hull` (313 bins).
# There is no .caffemodel parser offline (no Caffe, no caffe.proto): `caffemodel_path` is a torch state_dict file
# holding the Caffe blobs under the reference state_dict key names (+ the `caffe.*` keys of include/idc_b200.h), or an
# in-memory `state_dict`.  Spec-only: no Caffe weights or vectors exist offline (parity unpinned, oracle/caffe_spec.py).
# =============================================================================================
def caffe_scaled_state_dict(state_dict):
    """Caffe-scaled weights (conv1_1 trained on raw L-50 / ab / mask*110) -> the engine's PyTorch-scaled convention.
    Exact: w' . (L/100, ab/110, mask) == w . (L, ab, mask*110) with w' = w * (100, 110, 110, 110) per input channel."""
    import torch
    sd = dict(state_dict)
    w = sd["model1.0.weight"]
    w = w.detach().cpu().numpy() if hasattr(w, "detach") else np.asarray(w)
    scale = np.array([100.0, 110.0, 110.0, 110.0], dtype=np.float64).reshape(1, 4, 1, 1)
    sd["model1.0.weight"] = torch.from_numpy((w.astype(np.float64) * scale).astype(np.float32))
    return sd


class ColorizeImageB200Caffe(ColorizeImageB200):
    """<-> ColorizeImageCaffe (reference :375-442): regression model with the Caffe input / output scaling."""
    _caffe313 = False
    _global_hints = False

    def __init__(self, Xd=256, engine="tcgen05"):
        ColorizeImageB200.__init__(self, Xd, maskcent=False, engine=engine)
        self.mask_mult = 110.                     # reference :383
        self.pred_ab_layer = 'pred_ab'
        from . import prepost
        self.pts_in_hull = prepost.pts_in_hull().astype(np.float64)       # 313 x 2, in-gamut (reference :388-389)

    def prep_net(self, gpu_id=0, prototxt_path='', caffemodel_path='', state_dict=None):
        import torch
        from .engine import LhnContext
        print('gpu_id = %d, net_path = %s, model_path = %s' % (-1 if gpu_id is None else gpu_id, prototxt_path, caffemodel_path))
        if state_dict is None:
            state_dict = torch.load(caffemodel_path, map_location='cpu')
        self.gpu_id = gpu_id
        sd = caffe_scaled_state_dict(state_dict)
        if self._caffe313:
            sd["caffe.pts_in_hull"] = torch.from_numpy(self.pts_in_hull.astype(np.float32))   # reference :405-407
        self._ctx = LhnContext(device=0 if gpu_id in (None, -1) else int(gpu_id), max_n=1, H=self.Xd, W=self.Xd,
                               engine=self.engine, global_hints=self._global_hints, caffe313=self._caffe313,
                               options={"tanh_scale": 100})
        self._ctx.load_state_dict(sd)
        self.net_set = True

    def _engine_inputs(self):
        A = np.ascontiguousarray(self.img_l_mc, dtype=np.float32)[None]
        B = np.ascontiguousarray(self.input_ab_mc, dtype=np.float32)[None]
        M = np.ascontiguousarray(self.input_mask_mult / self.mask_mult, dtype=np.float32)[None]   # x110 lives in the weights
        return A, B, M

    def net_forward(self, input_ab, input_mask):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        self._click(self._ctx, 0.0, mask_div=self.mask_mult)     # the x110 of the mask lives in the conv1_1 weights
        return self.output_rgb                                    # output_ab_raw = the `pred_ab` blob (tanh * 100)


class ColorizeImageB200CaffeGlobDist(ColorizeImageB200Caffe):
    """<-> ColorizeImageCaffeGlobDist (reference :445-463): additional 313-bin global histogram input."""
    _global_hints = True

    def __init__(self, Xd=256, engine="tcgen05"):
        ColorizeImageB200Caffe.__init__(self, Xd, engine=engine)
        self.glob_mask_mult = 1.
        self.glob_layer = 'glob_ab_313_mask'

    def get_global_histogram(self, ref_rgb_u8):
        """DemoGlobalHistogramTransfer.ipynb:176-182 (gt_glob_net = global_stats.prototxt on the resized reference image)."""
        import cv2
        from . import prepost
        self.glob_vec = prepost.global_stats_gpu(cv2.resize(ref_rgb_u8, (self.Xd, self.Xd)), self._ctx.device)
        return self.glob_vec[:313].copy()

    def net_forward(self, input_ab, input_mask, glob_dist=-1):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        glob = np.zeros((1, 316), np.float32)            # "run without this, zero it out" (reference :454-456)
        if np.array(glob_dist).flatten()[0] != -1:
            glob[0, :313] = np.asarray(glob_dist, dtype=np.float32)
            glob[0, 313] = self.glob_mask_mult             # reference :458-459
        self._click(self._ctx, 0.0, glob=glob, mask_div=self.mask_mult)
        return self.output_rgb


class _LazyDist313(object):
    """[313, X, X] view of `dist_ab_S`: one pixel (313 floats) is computed on demand from the resident 313-bin logits
    (idc_caffe313_dist_pixel); the reference materialises 313 x X x X floats per forward and reads one pixel of it per
    click (reference :505, :521)."""

    def __init__(self, ctx, X, S):
        self._ctx, self._S = ctx, S
        self.shape = (313, X, X)
        self.dtype = np.dtype(np.float32)

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 3 and all(isinstance(i, (int, np.integer)) for i in idx[1:]):
            return self._ctx.caffe313_dist_pixel(0, int(idx[1]), int(idx[2]), self._S)[idx[0]]
        return self.__array__()[idx]

    def __array__(self, dtype=None, copy=None):
        X = self.shape[1]
        a = np.stack([np.stack([self._ctx.caffe313_dist_pixel(0, y, x, self._S) for x in range(X)], -1) for y in range(X)], -2)
        return a.astype(dtype) if dtype is not None else a


class ColorizeImageB200CaffeDist(ColorizeImageB200Caffe):
    """<-> ColorizeImageCaffeDist (reference :466-561): the 313-bin distribution model.  `pred_ab` is the annealed
    mean of the 313-bin head (deploy_nopred.prototxt:827-850), `dist_ab` the S-softened distribution (:808-820)."""
    _caffe313 = True

    def __init__(self, Xd=256, engine="tcgen05"):
        ColorizeImageB200Caffe.__init__(self, Xd, engine=engine)
        self.dist_ab_set = False
        self.scale_S_layer = 'scale_S'
        self.dist_ab_S_layer = 'dist_ab_S'
        g = np.arange(-110, 120, 10)
        # pts_grid.npy is (a, b)-ordered with a slowest: pts_grid[i] = (g[i // 23], g[i % 23])
        self.pts_grid = np.stack([np.repeat(g, 23), np.tile(g, 23)], axis=1)
        hull = set(map(tuple, self.pts_in_hull.astype(int).tolist()))
        self.in_hull = np.array([tuple(p) in hull for p in self.pts_grid.tolist()])      # identity: pts_grid[in_hull] == pts_in_hull
        self.AB = self.pts_grid.shape[0]
        self.A = self.B = int(np.sqrt(self.AB))
        self.dist_entropy = np.zeros((self.Xd, self.Xd))

    def prep_net(self, gpu_id=0, prototxt_path='', caffemodel_path='', S=.2, state_dict=None):
        ColorizeImageB200Caffe.prep_net(self, gpu_id, prototxt_path=prototxt_path, caffemodel_path=caffemodel_path,
                                        state_dict=state_dict)
        self.S = S

    def net_forward(self, input_ab, input_mask):
        if ColorizeImageBase.net_forward(self, input_ab, input_mask) == -1:
            return -1
        import torch
        from . import _lib
        A, B, M = self._engine_inputs()
        dev = "cuda:%d" % self._ctx.device
        dA, dB, dM = (torch.from_numpy(a).to(dev) for a in (A, B, M))
        self._ctx.forward_device(dA, dB, dM, 0.0)                      # trunk + hyper-column + pred_313 logits
        pred = self._ctx.caffe313_pred_ab(1, T=2.6)                    # annealed-mean `pred_ab` [1,2,X,X] (device)
        rgb = torch.empty((1, self.Xd, self.Xd, 3), dtype=torch.uint8, device=dev)
        L = (dA + 50.0).contiguous()
        st = torch.cuda.current_stream(self._ctx.device).cuda_stream
        rc = _lib.load().idc_lab2rgb_u8(self._ctx.device, 1, self.Xd, self.Xd, L.data_ptr(), pred.data_ptr(), rgb.data_ptr(), st)
        if rc != _lib.IDC_OK:
            raise _lib.IdcError(rc, "idc_lab2rgb_u8 failed")
        self.output_ab_raw = pred[0].cpu().numpy()
        self.output_rgb = rgb[0].cpu().numpy()
        self._set_out_ab_()
        self.dist_ab = _LazyDist313(self._ctx, self.Xd, self.S)
        self.dist_ab_set = True
        return self.output_rgb

    @property
    def dist_ab_full(self):
        full = np.zeros((self.AB, self.Xd, self.Xd))
        full[self.in_hull, :, :] = np.asarray(self.dist_ab)
        return full

    @property
    def dist_ab_grid(self):
        return self.dist_ab_full.reshape((self.A, self.B, self.Xd, self.Xd))

    def get_ab_reccs(self, h, w, K=5, N=25000, return_conf=False, method='gpu'):
        """reference :515-547 on the 313 in-gamut bins.  method='gpu': weighted k-means on the device (the N -> infinity
        limit, as ColorizeImageB200Dist); method='sampled': the reference's np.random + sklearn procedure."""
        if not self.dist_ab_set:
            print('Need to set prediction first')
            return 0
        pmf = np.asarray(self.dist_ab[:, int(h), int(w)], dtype=np.float64)
        if method == 'gpu':
            from .prepost import ab_reccs_pmf_gpu
            p529, q529 = np.zeros(529, np.float32), np.zeros((529, 2), np.float32)     # the kernel clusters 529 slots;
            p529[:313], q529[:313] = pmf, self.pts_in_hull                              # zero-weight padding is inert
            centers, conf, _ = ab_reccs_pmf_gpu(p529, K=K, pts=q529, device=self._ctx.device)
            centers, conf = centers.astype(np.float64), conf.astype(np.float64)
            return (centers, conf) if return_conf else centers
        if method != 'sampled':
            raise ValueError("method must be 'gpu' or 'sampled'")
        from sklearn.cluster import KMeans
        cmf = np.cumsum(pmf)
        cmf /= cmf[-1]
        samples = self.pts_in_hull[np.digitize(np.random.uniform(low=0, high=1.0, size=N), bins=cmf), :]
        km = KMeans(n_clusters=K).fit(samples)
        cnt = np.histogram(km.labels_, np.arange(0, K + 1))[0]
        order = np.argsort(cnt, axis=0)[::-1]
        centers, conf = km.cluster_centers_[order, :], 1. * cnt[order] / N
        return (centers, conf) if return_conf else centers

    def compute_entropy(self):
        d = np.asarray(self.dist_ab)
        self.dist_entropy = np.sum(d * np.log(d), axis=0)
