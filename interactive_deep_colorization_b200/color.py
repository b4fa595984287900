"""Host-side sRGB <-> CIE Lab (D65, 2 deg observer) used by the wrapper for image loading
(`load_image` / `set_image`) and `_set_out_ab_`.  Same conventions as scikit-image's
`color.rgb2lab / lab2rgb`, which the reference calls at
/root/reference/data/colorize_image.py:27,36,172,178 (scikit-image is not installed here).
Vectorised float64 numpy; the per-forward Lab->RGB post-process runs on the GPU instead
(csrc/idc_heads.cu lab2rgb_kernel)."""
import numpy as np

_M = np.array([[0.412453, 0.357580, 0.180423],
               [0.212671, 0.715160, 0.072169],
               [0.019334, 0.119193, 0.950227]], dtype=np.float64)
_MI = np.linalg.inv(_M)
_WHITE = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)
_EPS = 0.008856
_KAPPA = 7.787
_OFF = 16.0 / 116.0


def _to_unit(img):
    img = np.asarray(img)
    return img.astype(np.float64) / 255.0 if img.dtype == np.uint8 else img.astype(np.float64)


def rgb2lab(rgb):
    c = _to_unit(rgb)
    lin = np.where(c > 0.04045, ((np.maximum(c, 0.04045) + 0.055) / 1.055) ** 2.4, c / 12.92)
    t = (lin @ _M.T) / _WHITE
    f = np.where(t > _EPS, np.cbrt(np.maximum(t, _EPS)), _KAPPA * t + _OFF)
    fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]
    return np.stack([116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz)], axis=-1)


def lab2rgb(lab):
    lab = np.asarray(lab, dtype=np.float64)
    fy = (lab[..., 0] + 16.0) / 116.0
    fx = lab[..., 1] / 500.0 + fy
    fz = np.maximum(fy - lab[..., 2] / 200.0, 0.0)
    f = np.stack([fx, fy, fz], axis=-1)
    t = np.where(f > 0.2068966, f ** 3, (f - _OFF) / _KAPPA) * _WHITE
    lin = t @ _MI.T
    return np.where(lin > 0.0031308, 1.055 * np.maximum(lin, 0.0031308) ** (1.0 / 2.4) - 0.055, lin * 12.92)


def lab2rgb_transpose(img_l, img_ab):
    """1xHxW, 2xHxW -> HxWx3 uint8 (clip, *255, truncate): data/colorize_image.py:20-28."""
    lab = np.concatenate((img_l, img_ab), axis=0).transpose((1, 2, 0))
    return (np.clip(lab2rgb(lab), 0, 1) * 255).astype("uint8")


def rgb2lab_transpose(img_rgb):
    """HxWx3 -> 3xHxW: data/colorize_image.py:31-36."""
    return rgb2lab(img_rgb).transpose((2, 0, 1))
