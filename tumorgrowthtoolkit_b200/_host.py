"""Host-side set-up and post-processing shared by the three solvers.

These few lines of arithmetic define the CONTRACT of solve() (which grid the loop sees,
how many steps of which dt it runs), not the hot path: resample, seed, step-count law,
crop / restore.  Each function cites the reference lines whose behaviour it reproduces.
"""
import math

import numpy as np
from scipy.ndimage import zoom as _ndimage_zoom


def _out_shape(shape, factor):
    factors = list(factor) if np.ndim(factor) else [factor] * len(shape)
    return tuple(int(round(n * f)) for n, f in zip(shape, factors))    # scipy's own rule


def _device_resample(a, out_shape, lo=None, virt_shape=None, device=0):
    from . import _ops
    return _ops.resample_linear(a, out_shape, lo, virt_shape, device)


def resample_linear(a, factor, device=0):
    """Trilinear resample with scipy.ndimage.zoom(order=1) semantics (FK/FK.py:150-151,
    233-238) on the GPU -- 80 % of the reference's solve() wall time is this call (SURVEY.md
    finding 5).  float64 arrays only; other dtypes (the boolean-mask quirk of SURVEY.md 7) keep
    scipy's dtype-specific rounding and go through scipy itself."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        return _ndimage_zoom(a, factor, order=1)
    out_shape = _out_shape(a.shape, factor)
    if out_shape == a.shape:
        return a                    # zoom(x, 1.0, order=1) is the exact identity; nothing downstream writes to it
    if a.ndim == 3:
        return _device_resample(a, out_shape, device=device)
    assert a.ndim == 4 and out_shape[3] == a.shape[3], "4-D arrays are resampled per channel"
    out = np.empty(out_shape, dtype=np.float64)
    for ch in range(a.shape[3]):
        out[..., ch] = _device_resample(np.ascontiguousarray(a[..., ch]), out_shape[:3], device=device)
    return out


class LowResGrid:
    """The grid the time loop runs on (FK/FK.py:154-173)."""

    def __init__(self, lowres_shape, fullres_shape, res_factor, dx_mm, dy_mm, dz_mm, pct):
        self.shape = tuple(int(s) for s in lowres_shape[:3])
        self.full_shape = tuple(int(s) for s in fullres_shape[:3])
        # zoom factor that maps the low-res result back to the input grid  (FK.py:158)
        self.up_factor = tuple(n / float(o) for n, o in zip(self.full_shape, self.shape))
        self.h = (dx_mm / res_factor, dy_mm / res_factor, dz_mm / res_factor)  # FK.py:165-167
        self.seed = tuple(int(p * n) for p, n in zip(pct, self.shape))         # FK.py:171-173

    def upsample(self, a, device=0):
        return np.array(resample_linear(a, self.up_factor, device=device))

    def output_spec(self, box):
        """(lo, virt_shape, out_shape) of restore_tumor + zoom back to the input grid, for the device-side output path."""
        return box[0], self.shape, _out_shape(self.shape, self.up_factor)

    def upsample_crop(self, crop_arr, box, device=0):
        """restore_tumor + zoom back to the input grid in one device pass (FK.py:229-238)."""
        return _device_resample(crop_arr, _out_shape(self.shape, self.up_factor), lo=box[0], virt_shape=self.shape,
                                device=device)


def seed_gaussian(grid, init_scale):
    """Initial tumour: the clipped Gaussian of FK/FK.py:92-105 centred on the seed voxel
    (FK.py:191-192): 250/(4*pi*5)^1.5 * exp(-r^2/20), r in mm/init_scale, <=0.1 -> 0, >1 -> 1.
    Values below the 0.1 cut-off (r > 5.7*init_scale mm) are zero by construction, so only a
    window around the seed is evaluated; inside it the arithmetic is the reference's."""
    (nx, ny, nz), h, seed = grid.shape, grid.h, grid.seed
    spread = 5.0
    peak = 250 / np.power(4 * np.pi * spread, 3 / 2)
    reach = np.sqrt(4 * spread * np.log(peak / 0.1)) * init_scale * 1.001 + 1e-9    # mm
    out = np.zeros((nx, ny, nz))
    lo = [max(0, int(np.floor(c - reach / d)) - 1) for c, d in zip(seed, h)]
    hi = [min(n, int(np.ceil(c + reach / d)) + 2) for c, d, n in zip(seed, h, (nx, ny, nz))]
    if any(a >= b for a, b in zip(lo, hi)):
        return out
    x = (np.arange(lo[0], hi[0]) - seed[0])[:, None, None] * h[0] / init_scale
    y = (np.arange(lo[1], hi[1]) - seed[1])[None, :, None] * h[1] / init_scale
    z = (np.arange(lo[2], hi[2]) - seed[2])[None, None, :] * h[2] / init_scale
    g = peak * np.exp(-(np.power(x, 2) + np.power(y, 2) + np.power(z, 2)) / (4 * spread))
    g = np.where(g > 0.1, g, 0)
    out[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = np.where(g > 1, np.float64(1), g)
    return out


def steps_and_dt(nt_float, stopping_time):
    """`Nt` is a float: dt = T/Nt but ceil(Nt) steps are run (FK/FK.py:185-187)."""
    dt = stopping_time / nt_float
    return int(np.ceil(nt_float)), dt


def record_schedule(nsteps, n_series):
    """Step indices after which a snapshot is taken (FK/FK.py:208)."""
    if n_series is None:
        return None
    return np.linspace(0, nsteps - 1, n_series, dtype=int)


def bounding_box(mask, margin, shape):
    """Bounding box of the true voxels of `mask`, grown by `margin` and clamped
    (FK/FK.py:62-66).  Returns (min_coords, max_coords) as integer arrays."""
    mask = np.asarray(mask)
    if mask.ndim != 3:                       # (the reference's argwhere form, any rank)
        idx = np.argwhere(mask)
        lo = np.maximum(idx.min(axis=0) - margin, 0)
        hi = np.minimum(idx.max(axis=0) + margin + 1, shape)
        return lo, hi
    # same box from the three axis projections: argwhere materialises every true voxel's index triple (36 ms on the
    # atlas), the projections are three passes over a byte array (4 ms)
    mask = mask.astype(bool, copy=False)
    zy = mask.any(axis=0)
    proj = (mask.any(axis=(1, 2)), zy.any(axis=1), zy.any(axis=0))
    if not proj[0].any():
        raise ValueError("zero-size array to reduction operation minimum which has no identity")      # what argwhere().min() raises
    first = np.array([int(np.argmax(q)) for q in proj])
    last = np.array([len(q) - 1 - int(np.argmax(q[::-1])) for q in proj])
    lo = np.maximum(first - margin, 0)
    hi = np.minimum(last + margin + 1, shape)
    return lo, hi


def any_positive_last(a):
    """np.max(a, axis=-1) > 0 for an array whose last axis is short (the three rgb channels, FK_DTI.py:169): channel by
    channel -- numpy's reduction over a strided axis of length 3 takes 0.3 s on the atlas, this 0.05 s.  NaN-free input."""
    a = np.asarray(a)
    out = a[..., 0] > 0
    for c in range(1, a.shape[-1]):
        out |= a[..., c] > 0
    return out


def crop(a, box):
    lo, hi = box
    return a[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]


def paste(shape, a, box):
    """Inverse of crop into a zero array (FK/FK.py:75-90)."""
    out = np.zeros(shape)
    lo, hi = box
    out[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = a
    return out


def check_tissue_inputs(gm, wm, pct):
    """Input validation of FK/FK.py:140-147 (raises AssertionError like the reference)."""
    assert isinstance(gm, np.ndarray), "sGM must be a numpy array"
    assert isinstance(wm, np.ndarray), "sWM must be a numpy array"
    assert gm.ndim == 3, "sGM must be a 3D numpy array"
    assert wm.ndim == 3, "sWM must be a 3D numpy array"
    assert gm.shape == wm.shape
    check_seed(pct)


def check_seed(pct):
    for name, v in zip(("NxT1_pct", "NyT1_pct", "NzT1_pct"), pct):
        assert 0 <= v <= 1, f"{name} must be between 0 and 1"


def is_finite_volume(v):
    return v is not None and not (isinstance(v, float) and math.isnan(v))
