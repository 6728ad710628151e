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


def dti_lower6_field(seg, seed=20261019, sigma=6.0):
    """SURVEY.md 8(d) input 3: synthetic (X,Y,Z,6) lower-triangular tensor volume
    [dxx, dxy, dyy, dxz, dyz, dzz] (the storage TumorGrowthToolkit/FK_DTI/tools.py:91-106 unpacks) from a tissue
    label volume (3 = WM, 2 = GM): isotropic 1e-3*I in GM, eigenvalues (1.7, 0.3, 0.3)e-3 along a smooth unit
    direction field normalize(gaussian_filter(rng.normal(size=(3,)+shape), sigma)) in WM, zero in CSF and
    background (what FK_DTI_example.ipynb cell 3 does with its mask).  Stands in for the FSL_HCP1065 tensor
    volume, which the reference repository does not ship."""
    from scipy.ndimage import gaussian_filter
    seg = np.asarray(seg)
    rng = np.random.default_rng(seed)
    e = np.stack([gaussian_filter(rng.normal(size=seg.shape), sigma) for _ in range(3)])
    nrm = np.sqrt((e * e).sum(axis=0))
    nrm[nrm == 0] = 1.0
    e /= nrm
    is_wm, is_gm = seg == 3, seg == 2
    out = np.zeros(seg.shape + (6,), dtype=np.float64)
    for c, (a, b) in enumerate(((0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (2, 2))):
        aniso = 1.4e-3 * e[a] * e[b] + (0.3e-3 if a == b else 0.0)
        out[..., c] = np.where(is_wm, aniso, 0.0)
        if a == b:
            out[..., c] += np.where(is_gm, 1.0e-3, 0.0)
    return out


def seed_inside(mask, pct):
    """Seed fractions for FK_DTI_Solver (NxT1_pct, ...): `pct` itself when int(pct*n) lands inside `mask`, else the
    fractions of the mask voxel nearest to it (SURVEY.md 8(d) input 3: "re-pick inside WM if outside")."""
    shape = np.array(mask.shape)
    idx = tuple(int(p * n) for p, n in zip(pct, shape))
    if mask[idx]:
        return tuple(float(p) for p in pct)
    cand = np.argwhere(mask)
    best = cand[np.argmin(((cand - np.array(idx)) ** 2).sum(axis=1))]
    return tuple(float((b + 0.5) / n) for b, n in zip(best, shape))
