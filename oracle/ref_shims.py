"""Import the UNMODIFIED reference (/root/reference) in the build container.

The reference wrapper needs matplotlib and scikit-image, both absent here
(data/colorize_image.py:3-4).  We inject a stub ``matplotlib`` and a ``skimage.color``
backed by oracle/color_ref.py into sys.modules, then import the reference modules as
they are.  /root/reference does not exist on the GPU box: callers must check
``reference_available()`` and skip.

Test infrastructure only -- see oracle/__init__.py.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("IDC_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "pytorch", "model.py"))


def _install_shims():
    from . import color_ref
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "skimage" not in sys.modules:
        try:
            import skimage  # noqa: F401
        except ImportError:
            sk = types.ModuleType("skimage")
            col = types.ModuleType("skimage.color")
            col.rgb2lab = color_ref.rgb2lab
            col.lab2rgb = color_ref.lab2rgb
            sk.color = col
            sys.modules["skimage"] = sk
            sys.modules["skimage.color"] = col


def import_reference_model():
    """-> module /root/reference/models/pytorch/model.py (SIGGRAPHGenerator)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    return importlib.import_module("models.pytorch.model")


def import_reference_wrapper():
    """-> module /root/reference/data/colorize_image.py (ColorizeImageTorch...)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("data.colorize_image")
