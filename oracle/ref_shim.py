"""TEST INFRASTRUCTURE — import shim that lets the *unmodified* reference package
(`/root/reference/jolideco`) be imported in the build container.

It is only used by `oracle/make_golden.py` (to generate `tests/golden/*.npz`) and by the
oracle-validation tests that are skipped when `/root/reference` is absent (the GPU box has no
copy of the reference).  Nothing under `jolideco_b200/` may import this module.

The reference does all of its arithmetic with torch / numpy / scipy, but its modules import
`astropy`, `matplotlib` and a generated `jolideco.version` at top level, and read a GMM library
index at import time (`jolideco/priors/patches/gmm.py:493-508`).  None of those are installed
here, so minimal stand-ins are placed in `sys.modules` first.  The stand-ins restate published
behaviour of astropy (`Gaussian2DKernel`, `Tophat2DKernel`, `convolve`, `convolve_fft`,
`Table`, `lazyproperty`); they reproduce the reference's own data fixtures
(`jolideco/data/tests/test_core.py:16-43`) and its e2e goldens (`jolideco/tests/test_core.py:71-188`).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("JOLIDECO_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "jolideco"))


class _LazyProperty:
    """astropy.utils.lazyproperty: non-data descriptor caching into the instance dict."""

    def __init__(self, fget):
        self.fget = fget
        self.__doc__ = fget.__doc__
        self._key = fget.__name__

    def __get__(self, obj, owner=None):
        if obj is None:
            return self
        val = self.fget(obj)
        obj.__dict__[self._key] = val
        return val


class _Row(dict):
    pass


class _Table:
    """astropy.table.Table subset used by `jolideco/loss.py:192-250` and `core.py:245-267`."""

    def __init__(self, names=None, dtype=None, **kwargs):
        self.colnames = list(names or [])
        self.dtype = list(dtype or [])
        self.rows = []
        self.meta = {}

    def add_row(self, row):
        self.rows.append(_Row(row))

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, item):
        if isinstance(item, str):
            return np.array([r[item] for r in self.rows])
        if isinstance(item, slice):
            t = _Table(self.colnames, self.dtype)
            t.rows = self.rows[item]
            return t
        return self.rows[item]

    def __setitem__(self, key, value):
        self.colnames.append(key)


def _kernel_coords(size, oversample):
    # astropy.convolution.utils.discretize_model: pixel centres at integers for odd sizes,
    # half-integers for even sizes; "oversample" = mean over a factor x factor sub-grid.
    if size % 2:
        lo, hi = -(size - 1) / 2.0, (size - 1) / 2.0 + 1
    else:
        lo, hi = -size / 2.0 + 0.5, size / 2.0 + 0.5
    if oversample is None:
        return np.arange(lo, hi)
    f = oversample
    return np.linspace(lo - 0.5 * (1 - 1.0 / f), hi - 0.5 * (1 + 1.0 / f), num=int((hi - lo) * f))


class _Kernel2D:
    def __init__(self, func, default_size, x_size=None, y_size=None, mode="center", factor=10):
        x_size = default_size if x_size is None else x_size
        y_size = x_size if y_size is None else y_size
        over = factor if mode == "oversample" else None
        x = _kernel_coords(x_size, over)
        y = _kernel_coords(y_size, over)
        xx, yy = np.meshgrid(x, y)
        vals = func(xx, yy)
        if over:
            vals = vals.reshape(y_size, over, x_size, over).mean(axis=(1, 3))
        self._array = vals / vals.sum()

    @property
    def array(self):
        return self._array

    def __array__(self, dtype=None, copy=None):
        return self._array if dtype is None else self._array.astype(dtype)


def _odd_size(x):
    n = int(np.ceil(x))
    return n if n % 2 else n + 1


class Gaussian2DKernel(_Kernel2D):
    def __init__(self, x_stddev, y_stddev=None, theta=0.0, **kwargs):
        s = float(x_stddev)

        def func(x, y):
            return np.exp(-0.5 * (x**2 + y**2) / s**2) / (2 * np.pi * s**2)

        super().__init__(func, _odd_size(8 * s), **kwargs)


class Tophat2DKernel(_Kernel2D):
    def __init__(self, radius, **kwargs):
        r = float(radius)

        def func(x, y):
            return ((x**2 + y**2) <= r**2) / (np.pi * r**2)

        super().__init__(func, _odd_size(2 * r), **kwargs)


def _convolve(array, kernel, **kwargs):
    from scipy.signal import fftconvolve

    k = np.asarray(kernel.array if hasattr(kernel, "array") else kernel, dtype=float)
    return fftconvolve(np.asarray(array, dtype=float), k / k.sum(), mode="same")


def install():
    """Install the stand-in modules and put the reference on sys.path. Idempotent."""
    if "jolideco" in sys.modules and getattr(sys.modules["jolideco"], "_shimmed", False):
        return sys.modules["jolideco"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "astropy" not in sys.modules:
        astropy = mod("astropy")
        astropy.utils = mod("astropy.utils", lazyproperty=_LazyProperty)
        astropy.table = mod("astropy.table", Table=_Table)
        astropy.convolution = mod(
            "astropy.convolution",
            Gaussian2DKernel=Gaussian2DKernel,
            Tophat2DKernel=Tophat2DKernel,
            convolve=_convolve,
            convolve_fft=_convolve,
        )
        astropy.visualization = mod("astropy.visualization", simple_norm=lambda *a, **k: None)
        astropy.coordinates = mod("astropy.coordinates", SkyCoord=object)
        astropy.wcs = mod("astropy.wcs", WCS=object)
        astropy.io = mod("astropy.io")
        astropy.io.fits = mod("astropy.io.fits")
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
    mod("jolideco.version", version="0.3.dev0+shim")

    if "JOLIDECO_GMM_LIBRARY" not in os.environ:
        d = tempfile.mkdtemp(prefix="jolideco-gmm-index-")
        with open(os.path.join(d, "jolideco-gmm-library-index.json"), "w") as fh:
            json.dump({}, fh)
        os.environ["JOLIDECO_GMM_LIBRARY"] = d

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import jolideco  # noqa: E402

    jolideco._shimmed = True
    # the reference wraps FluxComponents in torch.compile (core.py:183-184): a numerical
    # no-op that only costs compile time; disable for golden generation.
    import jolideco.core as jcore

    jcore.COMPILE_MODEL = False
    return jolideco
