"""CPU restatement of the reference's tensor elongation -- TEST INFRASTRUCTURE, never imported by the
product path.  Follows TumorGrowthToolkit/FK_DTI/tools.py:109-164 line by line with numpy's float32
LAPACK eigh in place of torch's: scale the largest eigenvalue, shift the other two, rebuild V diag(e) V^T.
Pinned against reference-generated vectors (tests/golden/elongate.npz, tests/make_golden.py)."""
import numpy as np


def elongate(tensors, scale_factor):
    t = np.asarray(tensors).astype(np.float32)                       # tools.py:114 (.float())
    e, v = np.linalg.eigh(t)                                         # :117, ascending eigenvalues
    total = e.sum(axis=-1, keepdims=True)                            # :119
    imax = np.argmax(e, axis=-1)[..., None]                          # :121, first index on ties
    emax = np.take_along_axis(e, imax, -1)                           # :122
    scaled = (emax * np.float32(scale_factor)).astype(np.float32)    # :123
    adjustment = (scaled - emax) / np.float32(2)                     # :126-129
    others = np.ones(e.shape, dtype=bool)
    np.put_along_axis(others, imax, False, -1)                       # :130-131
    e_adj = np.where(others, e - adjustment, e)                      # :134
    fix = (total - e_adj.sum(axis=-1, keepdims=True)) / np.float32(3)    # :135-138
    e_fin = e_adj + np.where(others, fix, np.float32(0))             # :139
    np.put_along_axis(e_fin, imax, scaled, -1)                       # :142
    return np.einsum("...ij,...j,...kj->...ik", v, e_fin, v).astype(np.float32)   # :145
