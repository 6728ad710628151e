"""DTI input preparation used by FK_DTI_Solver.solve (reference: TumorGrowthToolkit/FK_DTI/tools.py).

Runs once per solve, before the time loop.  Unlike the reference module this one imports
neither dipy, nibabel nor matplotlib, so `import tumorgrowthtoolkit_b200.FK_DTI` always works.
"""
import numpy as np


def get_tensor_from_lower6(lower6):
    """(X,Y,Z,6) lower-triangular storage [dxx, dxy, dyy, dxz, dyz, dzz] -> (X,Y,Z,3,3)
    symmetric tensors (tools.py:91-106)."""
    lower6 = np.asarray(lower6)
    t = np.zeros(lower6.shape[0:3] + (3, 3))
    print(t.shape)                                   # tools.py:95 prints the shape
    for (a, b), c in (((0, 0), 0), ((1, 1), 2), ((2, 2), 5), ((0, 1), 1), ((0, 2), 3), ((1, 2), 4)):
        t[..., a, b] = lower6[..., c]
        t[..., b, a] = lower6[..., c]
    return t


def elongate_tensor_along_main_axis_torch(tensor_arrayNP, scale_factor):
    """Scale the largest eigenvalue by `scale_factor`, shrink the other two so the trace is
    kept, rebuild the tensors; float32 torch.linalg.eigh on the CPU like the reference
    (tools.py:109-164).  One-off preprocessing before the time loop.  This is the literal restatement
    (bit-identical to the reference on the same torch build) that FK_DTI_Solver uses by default;
    `elongate_tensor_along_main_axis_cuda` is the GPU kernel (params['elongation_backend'] = 'cuda')."""
    import torch
    with torch.no_grad():
        dev = "cpu"
        t = torch.from_numpy(np.ascontiguousarray(tensor_arrayNP)).float().to(dev)
        e, v = torch.linalg.eigh(t)
        total = e.sum(dim=-1, keepdim=True)
        imax = torch.argmax(e, dim=-1, keepdim=True)
        emax = torch.gather(e, -1, imax)
        emax_scaled = emax * scale_factor
        others = torch.ones_like(e, dtype=torch.bool)
        others.scatter_(-1, imax, False)
        e_adj = torch.where(others, e - (emax_scaled - emax) / 2, e)
        fix = (total - e_adj.sum(dim=-1, keepdim=True)) / 3
        e_fin = e_adj + torch.where(others, fix, torch.zeros_like(fix))
        e_fin.scatter_(-1, imax, emax_scaled)
        out = v @ torch.diag_embed(e_fin) @ v.transpose(-2, -1)
        return out.cpu().numpy()


def elongate_tensor_along_main_axis_cuda(tensor_arrayNP, scale_factor, device=0):
    """The same elongation as a hand-written CUDA kernel (csrc/tgtk_elongate.cuh):
    float64 closed-form top eigenpair per tensor instead of float32 LAPACK, float32 result like the
    reference's.  Agrees with `elongate_tensor_along_main_axis_torch` to float32 rounding (1e-6 of the
    largest entry), including the axis LAPACK picks for exactly diagonal degenerate tensors."""
    from .. import _ops
    return _ops.elongate_tensors(tensor_arrayNP, scale_factor, device=device)


def makeXYZ_rgb_from_tensor(tensor, exponent=1, linear=0):
    """Tensor diagonal -> per-axis diffusivity weights: brain-mean 1, clipped to [0,10],
    x**exponent + linear*x, renormalised and clipped again (tools.py:166-187)."""
    out = np.zeros(tensor.shape[:4])
    for a in range(3):
        out[:, :, :, a] = tensor[:, :, :, a, a]
    brain = np.max(out, axis=-1) > 0

    def normalise(x):
        x /= np.mean(x[brain])
        x[x > 10] = 10
        x[x < 0] = 0
        return x

    out = normalise(out)
    out = out ** exponent + out * linear
    return normalise(out)


def makeXYZ_rgb_from_tensor_cuda(tensor, exponent=1, linear=0, device=0):
    """`makeXYZ_rgb_from_tensor` as device kernels (csrc: rgb_extract / rgb_mean / rgb_normalise): what FK_DTI_Solver.solve
    runs on the whole input volume before the loop.  Differs from the numpy function above only in the summation order
    of the two brain means (~1e-15 relative)."""
    from .. import _ops
    return _ops.dti_rgb(tensor, exponent, linear, device=device)


def offdiagonal_from_tensor(tensor):
    """(X,Y,Z,3) array [D_xy, D_xz, D_yz] in the units of makeXYZ_rgb_from_tensor(tensor, 1, 0):
    the same two brain-mean normalisations, no clipping (|D_ab| <= sqrt(D_aa D_bb) anyway).
    Only used by the opt-in `dti_full_tensor` mode, which the reference does not have."""
    diag = np.stack([tensor[..., a, a] for a in range(3)], axis=-1).astype(np.float64)
    brain = np.max(diag, axis=-1) > 0
    m1 = np.mean(diag[brain])
    d1 = np.clip(diag / m1, 0, 10)
    m2 = np.mean(d1[brain])
    off = np.stack([tensor[..., 0, 1], tensor[..., 0, 2], tensor[..., 1, 2]], axis=-1).astype(np.float64)
    return off / (m1 * m2)


def get_RGB(*args, **kwargs):
    """Raw-DWI -> tensor fit with dipy (tools.py:26-89): outside the time-stepping path and
    dipy is not a dependency of this package."""
    raise NotImplementedError("get_RGB needs dipy/nibabel; fit the tensor with the reference toolkit "
                              "and pass `diffusionTensors` to FK_DTI_Solver")
