"""GPU versions of the steps either side of the network.

    rgb2lab_gpu        skimage color.rgb2lab on uint8 RGB  (reference data/colorize_image.py:31-36,172-178,196-198)
    fullres_rgb_gpu    get_img_fullres (:123-131): scipy zoom(order=1) of ab + Lab->RGB at full resolution

float64 arithmetic on the device, like the reference's numpy path.  torch supplies device memory only.
"""
import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def rgb2lab_gpu(rgb_u8, device=0):
    """HxWx3 (or NxHxWx3) uint8 -> 3xHxW (or Nx3xHxW) float64 numpy, == color.rgb2lab(...).transpose."""
    torch = _torch()
    a = np.ascontiguousarray(rgb_u8)
    assert a.dtype == np.uint8 and a.shape[-1] == 3
    single = a.ndim == 3
    if single:
        a = a[None]
    n, h, w = a.shape[:3]
    d_rgb = torch.from_numpy(a).to("cuda:%d" % device)
    d_lab = torch.empty((n, 3, h, w), dtype=torch.float64, device=d_rgb.device)
    st = torch.cuda.current_stream(d_rgb.device).cuda_stream
    rc = _lib.load().idc_rgb2lab_f64(device, n, h, w, d_rgb.data_ptr(), d_lab.data_ptr(), st)
    if rc != _lib.IDC_OK:
        raise _lib.IdcError(rc, "idc_rgb2lab_f64 failed")
    out = d_lab.cpu().numpy()
    return out[0] if single else out


def fullres_rgb_gpu(ab, l_fullres, device=0):
    """ab [2,h,w] (any float), l_fullres [1,H,W] or [H,W] float64 -> uint8 [H,W,3]."""
    torch = _torch()
    ab = np.ascontiguousarray(ab, dtype=np.float64)
    L = np.ascontiguousarray(np.asarray(l_fullres, dtype=np.float64).reshape(l_fullres.shape[-2], l_fullres.shape[-1]))
    H, W = L.shape
    d_ab = torch.from_numpy(ab).to("cuda:%d" % device)
    d_L = torch.from_numpy(L).to(d_ab.device)
    d_rgb = torch.empty((H, W, 3), dtype=torch.uint8, device=d_ab.device)
    st = torch.cuda.current_stream(d_ab.device).cuda_stream
    rc = _lib.load().idc_zoom_lab2rgb_u8(device, ab.shape[1], ab.shape[2], d_ab.data_ptr(), H, W, d_L.data_ptr(),
                                         d_rgb.data_ptr(), st)
    if rc != _lib.IDC_OK:
        raise _lib.IdcError(rc, "idc_zoom_lab2rgb_u8 failed")
    return d_rgb.cpu().numpy()


def pts_in_hull():
    """The 313 in-gamut ab bin centres (data fixture of the reference: data/color_bins/pts_in_hull.npy)."""
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pts_in_hull.npy")).astype(np.float32)


def global_stats_gpu(rgb_u8, device=0):
    """Reference image uint8 [H,W,3] (H, W multiples of 4) -> glob vector [316] =
    [313-bin ab histogram, 1, mean saturation, 1] (row f3; global_stats.prototxt)."""
    torch = _torch()
    a = np.ascontiguousarray(rgb_u8)
    assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3
    d_rgb = torch.from_numpy(a).to("cuda:%d" % device)
    d_pts = torch.from_numpy(pts_in_hull()).to(d_rgb.device)
    d_out = torch.empty((316,), dtype=torch.float32, device=d_rgb.device)
    st = torch.cuda.current_stream(d_rgb.device).cuda_stream
    rc = _lib.load().idc_global_stats(device, a.shape[0], a.shape[1], d_rgb.data_ptr(), d_pts.data_ptr(), d_out.data_ptr(), st)
    if rc != _lib.IDC_OK:
        raise _lib.IdcError(rc, "idc_global_stats failed")
    return d_out.cpu().numpy()


def ab_reccs_pmf_gpu(pmf, K=5, max_iter=100, n_init=8, pts=None, device=0):
    """Colour suggestions for one 529-bin pmf (host array): the deterministic weighted-k-means form of the
    reference's get_ab_reccs (data/colorize_image.py:322-354), see include/idc_b200.h: idc_ab_reccs_pmf.
    Returns (centres [K,2], mass [K], Lloyd iterations)."""
    import ctypes
    p = np.ascontiguousarray(pmf, np.float32)
    assert p.shape == (529,)
    q = None if pts is None else np.ascontiguousarray(pts, np.float32)
    assert q is None or q.shape == (529, 2)
    centers, conf, iters = np.empty((K, 2), np.float32), np.empty((K,), np.float32), ctypes.c_int(0)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = _lib.load().idc_ab_reccs_pmf(device, vp(p), int(K), int(max_iter), int(n_init), None if q is None else vp(q),
                                      vp(centers), vp(conf), ctypes.byref(iters))
    if rc != _lib.IDC_OK:
        raise _lib.IdcError(rc, "idc_ab_reccs_pmf failed")
    return centers, conf, iters.value
