"""The reference's public per-step methods served by the CUDA library, one call each
(upload -> kernel -> download).  solve() does not come through here: its loop stays on the
device (see _loop.py); these exist because callers and subclasses of the reference invoke
FK_update / get_D / FK_2c_update / get_D_with_necro / m_Tildas directly
(FK_DTI/FK_DTI.py:196, notebooks, scripts).
"""
import ctypes

import numpy as np

from . import _native as nat

_KEYS = ("D_plus_x", "D_plus_y", "D_plus_z", "D_minus_x", "D_minus_y", "D_minus_z")
_I64x3 = ctypes.c_int64 * 3
_I64x6x3 = (ctypes.c_int64 * 3) * 6
_PTR6 = ctypes.c_void_p * 6


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _faces(WM, GM, th, Dw, ratio, P, N, th_necro, what, device):
    L = nat.lib()
    logical_or = int(np.asarray(WM).dtype == np.bool_ and np.asarray(GM).dtype == np.bool_)
    wm, gm = nat._as_f64(WM), nat._as_f64(GM)
    assert wm.ndim == 3 and wm.shape == gm.shape
    outs = [np.empty(wm.shape, dtype=np.float64) for _ in range(3)]
    if P is not None:
        pc, nc = nat._as_f64(P), nat._as_f64(N)
        pargs = (_vp(pc), nat._strides(pc), _vp(nc), nat._strides(nc))
    else:
        pargs = (None, _I64x3(), None, _I64x3())
    rc = L.tgtk_faces_host(int(device), *[ctypes.c_int(s) for s in wm.shape], _vp(wm), nat._strides(wm),
                           _vp(gm), nat._strides(gm), *pargs, ctypes.c_double(th), ctypes.c_double(th_necro),
                           ctypes.c_double(Dw), ctypes.c_double(ratio), ctypes.c_int(logical_or),
                           ctypes.c_int(what), *[_vp(o) for o in outs])
    nat._check(rc)
    return outs


def face_coefficients(WM, GM, th, Dw, ratio, P=None, N=None, th_necro=0.0, device=0):
    """get_D (FK/FK.py:22-32) / get_D_with_necro (FK_2c/FK_2c.py:36-46): dict of six arrays.
    D_plus_a is D_minus_a shifted one voxel up with wrap (np.roll(...,1,axis=a), FK.py:28-30)."""
    m = _faces(WM, GM, th, Dw, ratio, P, N, th_necro, 0, device)
    out = {}
    for a, name in enumerate("xyz"):
        out["D_minus_" + name] = m[a]
        out["D_plus_" + name] = np.roll(m[a], 1, axis=a)
    return out


def tissue_tildas(WM, GM, th, P=None, N=None, th_necro=0.0, device=0):
    """m_Tildas (FK/FK.py:10-20) / m_Tildas_with_necro (FK_2c/FK_2c.py:10-34)."""
    w = _faces(WM, GM, th, 1.0, 1.0, P, N, th_necro, 1, device)
    g = _faces(WM, GM, th, 1.0, 1.0, P, N, th_necro, 2, device)
    return {"WM_t_x": w[0], "WM_t_y": w[1], "WM_t_z": w[2], "GM_t_x": g[0], "GM_t_y": g[1], "GM_t_z": g[2]}


def _coef_args(D):
    arrs = [nat._as_f64(D[k]) for k in _KEYS]
    ptrs = _PTR6(*[a.ctypes.data for a in arrs])
    strides = _I64x6x3()
    for c, a in enumerate(arrs):
        for d in range(3):
            strides[c][d] = a.strides[d] // a.itemsize
    return arrs, ptrs, strides


def _inplace(a):
    """float64 array to hand to the library plus a flag telling whether it IS `a`."""
    arr = np.asarray(a)
    if arr.dtype == np.float64:
        return arr, True
    return arr.astype(np.float64), False


def fk_update(A, D, rho, dt, dx, dy, dz, device=0, precision="fp64"):
    """FK_update (FK/FK.py:34-42): A is updated in place and returned."""
    from ._loop import dtype_code
    L = nat.lib()
    buf, same = _inplace(A)
    p = nat.make_params(nat.MODEL_COEF6, dtype_code(precision), buf.shape, (dx, dy, dz), dt, rho=rho)
    arrs, ptrs, strides = _coef_args(D)
    for a in arrs:
        if a.shape != buf.shape:
            raise ValueError("coefficient arrays must have the shape of A")
    nat._check(L.tgtk_fk_update_host(ctypes.byref(p), int(device), _vp(buf), nat._strides(buf), ptrs, strides))
    if not same:
        A[...] = buf
    return A


def fk2c_update(P, N, S, D, rho, lambda_np, sigma_np, Ds, lambda_s, dt, dx, dy, dz, device=0, precision="fp64"):
    """FK_2c_update (FK_2c/FK_2c.py:48-73): P, N, S are updated in place; returns the dict."""
    from ._loop import dtype_code
    L = nat.lib()
    bufs = [_inplace(x) for x in (P, N, S)]
    shape = bufs[0][0].shape
    p = nat.make_params(nat.MODEL_FK2C_COEF12, dtype_code(precision), shape, (dx, dy, dz), dt, rho=rho,
                        lambda_np=lambda_np, sigma_np=sigma_np, lambda_s=lambda_s)
    dkeep, dptrs, dstr = _coef_args(D)      # keep the float64 views alive across the call
    skeep, sptrs, sstr = _coef_args(Ds)
    args = []
    for b, _same in bufs:
        args += [_vp(b), nat._strides(b)]
    nat._check(L.tgtk_fk2c_update_host(ctypes.byref(p), int(device), *args, dptrs, dstr, sptrs, sstr))
    del dkeep, skeep
    for (b, same), dst in zip(bufs, (P, N, S)):
        if not same:
            dst[...] = b
    return {'P': P, 'N': N, 'S': S}


_I32x3 = ctypes.c_int32 * 3


def resample_linear(a, out_shape, lo=None, virt_shape=None, device=0):
    """scipy.ndimage.zoom(v, f, order=1) on the GPU (FK/FK.py:150-151,233-238).  `a` (float64,
    any strides) fills the box starting at `lo` of a zero volume of shape `virt_shape`
    (restore_tumor, FK/FK.py:75-90, fused in); plain zoom when lo/virt_shape are omitted."""
    L = nat.lib()
    a = nat._as_f64(a)
    assert a.ndim == 3
    if virt_shape is None:
        virt_shape, lo = a.shape, (0, 0, 0)
    out = np.empty(tuple(int(v) for v in out_shape), dtype=np.float64)
    rc = L.tgtk_resample_linear_host(int(device), _vp(a), nat._strides(a), _I32x3(*a.shape),
                                     _I32x3(*[int(v) for v in lo]), _I32x3(*[int(v) for v in virt_shape]),
                                     _vp(out), _I32x3(*out.shape))
    nat._check(rc)
    return out


def elongate_tensors(tensors, scale_factor, device=0):
    """tools.elongate_tensor_along_main_axis_torch (TumorGrowthToolkit/FK_DTI/tools.py:109-164) on the GPU:
    (..., 3, 3) symmetric tensors -> float32 array of the same shape with the largest eigenvalue of every
    tensor multiplied by `scale_factor` and the other two shifted as the reference shifts them."""
    L = nat.lib()
    t = np.ascontiguousarray(tensors, dtype=np.float64)
    assert t.ndim >= 2 and t.shape[-2:] == (3, 3), "tensors must have shape (..., 3, 3)"
    out = np.empty(t.shape, dtype=np.float32)
    n = int(np.prod(t.shape[:-2], dtype=np.int64))
    nat._check(L.tgtk_elongate_tensors_host(int(device), t.ctypes.data, n, float(scale_factor), out.ctypes.data))
    return out


def dti_rgb(tensors, exponent=1, linear=0, device=0):
    """tools.makeXYZ_rgb_from_tensor (TumorGrowthToolkit/FK_DTI/tools.py:166-187) on the GPU: (X,Y,Z,3,3) tensors ->
    (X,Y,Z,3) float64 per-axis diffusivity weights (brain-mean 1, clipped to [0, 10], x**exponent + linear*x, renormalised)."""
    L = nat.lib()
    t = np.ascontiguousarray(tensors, dtype=np.float64)
    assert t.ndim == 5 and t.shape[-2:] == (3, 3), "tensors must have shape (X, Y, Z, 3, 3)"
    out = np.empty(t.shape[:3] + (3,), dtype=np.float64)
    nat._check(L.tgtk_dti_rgb_host(int(device), t.ctypes.data, int(np.prod(t.shape[:3], dtype=np.int64)), float(exponent), float(linear),
                                   out.ctypes.data))
    return out
