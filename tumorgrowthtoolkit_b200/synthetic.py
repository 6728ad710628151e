"""Deterministic synthetic inputs of the reference's shapes (no file access).

BASELINE.json allows "synthetic volumes of the same 240x240x155 shape" in place of the
bundled atlas; the reference's GM.nii.gz and DTI tensor volume are absent anyway
(SURVEY.md finding 4).  Everything here is plain numpy and seeded.
"""
import numpy as np


def brain_phantom(shape=(240, 240, 155), dtype=np.float64):
    """WM / GM partial-volume maps in [0,1] of a brain-like ellipsoid.

    WM core, GM cortex shell of ~12 % of the radius, two CSF ventricles, a smooth
    (partial-volume) transition of about one voxel at every interface, zero background
    with a margin so the cropped box never touches the volume edge.
    Returns (wm, gm).
    """
    nx, ny, nz = shape
    x = (np.arange(nx, dtype=np.float64) - (nx - 1) / 2.0) / (0.36 * nx)
    y = (np.arange(ny, dtype=np.float64) - (ny - 1) / 2.0) / (0.42 * ny)
    z = (np.arange(nz, dtype=np.float64) - (nz - 1) / 2.0) / (0.40 * nz)
    xx, yy, zz = np.meshgrid(x, y, z, indexing="ij", sparse=True)
    # gyri-like modulation of the outer surface
    wobble = 0.04 * np.sin(9.0 * xx) * np.cos(7.0 * yy) + 0.03 * np.sin(11.0 * zz + 2.0 * xx)
    r = np.sqrt(xx * xx + yy * yy + zz * zz) + wobble
    sharp = 0.5 * min(nx, ny, nz) * 0.36          # ~1 voxel wide transitions

    def smooth_step(v):
        return np.clip(0.5 + v * sharp, 0.0, 1.0)

    brain = smooth_step(1.0 - r)
    core = smooth_step(0.86 - r + 0.05 * np.sin(13.0 * yy) * np.sin(12.0 * zz))
    vent = np.zeros_like(r)
    for cx in (-0.16, 0.16):
        rv = np.sqrt(((xx - cx) / 0.10) ** 2 + ((yy + 0.05) / 0.34) ** 2 + ((zz - 0.05) / 0.16) ** 2)
        vent = np.maximum(vent, smooth_step((1.0 - rv) * 0.15))
    wm = core * (1.0 - vent)
    gm = (brain - core) * (1.0 - vent)
    return np.ascontiguousarray(wm, dtype=dtype), np.ascontiguousarray(np.clip(gm, 0.0, 1.0), dtype=dtype)


def tissue_from_labels(seg):
    """wm=(seg==3), gm=(seg==2) as float64 -- the construction the reference's
    FK_DTI_example.ipynb (cell 9) uses on the SRI label atlas, cast to float."""
    seg = np.asarray(seg)
    return (seg == 3).astype(np.float64), (seg == 2).astype(np.float64)


def _smooth_unit_field(shape, rng, passes=3, width=9):
    """Smooth random unit vector field (box-filter smoothing; no scipy needed)."""
    f = rng.normal(size=(3,) + tuple(shape))
    for _ in range(passes):
        for ax in (1, 2, 3):
            f = _boxfilter(f, ax, width)
    nrm = np.sqrt((f * f).sum(axis=0))
    nrm[nrm == 0] = 1.0
    return f / nrm


def _boxfilter(f, axis, width):
    c = np.cumsum(f, axis=axis)
    pad = [(0, 0)] * f.ndim
    pad[axis] = (1, 0)
    c = np.pad(c, pad)
    n = f.shape[axis]
    lo = np.clip(np.arange(n) - width // 2, 0, n)
    hi = np.clip(np.arange(n) + width // 2 + 1, 0, n)
    return (np.take(c, hi, axis=axis) - np.take(c, lo, axis=axis)) / width


def dti_tensor_field(wm, gm, seed=20261019):
    """Synthetic (X,Y,Z,3,3) diffusion tensors (SURVEY.md 8(d) input 3): isotropic 1e-3*I in
    GM, eigenvalues (1.7,0.3,0.3)e-3 along a smooth direction field in WM, zero elsewhere."""
    rng = np.random.default_rng(seed)
    shape = wm.shape
    e = _smooth_unit_field(shape, rng)
    T = np.zeros(shape + (3, 3), dtype=np.float64)
    is_wm = wm > 0.5
    is_gm = (gm > 0.5) & ~is_wm
    for a in range(3):
        for b in range(3):
            aniso = 0.3e-3 * (1.0 if a == b else 0.0) + 1.4e-3 * e[a] * e[b]
            T[..., a, b] = np.where(is_wm, aniso, 0.0)
        T[..., a, a] += np.where(is_gm, 1.0e-3, 0.0)
    return T
