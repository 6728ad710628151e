"""Tensor elongation (TumorGrowthToolkit/FK_DTI/tools.py:109-164): the numpy oracle and the
torch mirror against reference-generated vectors (CPU), the CUDA kernel against both (GPU)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SCALES = (2.0, 0.6, 1.25)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "elongate.npz"))


def test_inputs_are_the_committed_generator(gold):
    from tests import make_golden
    assert np.array_equal(make_golden.elongate_inputs(), gold["tensors"])


@pytest.mark.parametrize("scale", SCALES)
def test_oracle_matches_reference_vectors(gold, scale):
    from oracle.elongate import elongate
    got, ref = elongate(gold["tensors"], scale), gold["scale_%g" % scale]
    assert got.dtype == ref.dtype == np.float32
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))


@pytest.mark.parametrize("scale", SCALES)
def test_torch_mirror_matches_reference_vectors(gold, scale):
    from tumorgrowthtoolkit_b200.FK_DTI import tools
    got, ref = tools.elongate_tensor_along_main_axis_torch(gold["tensors"], scale), gold["scale_%g" % scale]
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 2e-6 * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("scale", SCALES)
def test_cuda_kernel_matches_reference_vectors(gold, scale):
    import __graft_entry__ as ge
    ge.build()
    from tumorgrowthtoolkit_b200.FK_DTI import tools
    got, ref = tools.elongate_tensor_along_main_axis_cuda(gold["tensors"], scale), gold["scale_%g" % scale]
    assert got.dtype == np.float32 and got.shape == ref.shape
    err = np.abs(got - ref).reshape(-1, 9).max(axis=-1)
    assert err.max() <= 2e-6 * np.max(np.abs(ref)), (int(np.argmax(err)), float(err.max()))


@pytest.mark.gpu
def test_cuda_kernel_large_batch_properties():
    """Size-independent properties on 300k random SPD tensors (several blocks, a ragged last one): symmetry,
    eigenvalues (e_max*s, e_i - (e_max*s - e_max)/6), eigenvectors unchanged, agreement with the numpy oracle."""
    from oracle.elongate import elongate
    from tumorgrowthtoolkit_b200 import _ops
    rng = np.random.default_rng(5)
    n, s = 300_007, 1.7
    q, _ = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    lam = np.sort(rng.uniform(0.2, 1.0, (n, 3)), axis=-1) * np.array([1.0, 1.5, 2.4])
    t = np.einsum("nij,nj,nkj->nik", q, lam, q)
    t = 0.5 * (t + np.swapaxes(t, -1, -2))
    got = _ops.elongate_tensors(t, s).astype(np.float64)
    assert np.max(np.abs(got - np.swapaxes(got, -1, -2))) <= 1e-6
    e = np.linalg.eigvalsh(got)
    want = np.stack([lam[:, 0] - (lam[:, 2] * s - lam[:, 2]) / 6, lam[:, 1] - (lam[:, 2] * s - lam[:, 2]) / 6, lam[:, 2] * s], -1)
    assert np.max(np.abs(e - np.sort(want, -1))) <= 5e-6
    ref = elongate(t, s)
    assert np.max(np.abs(got - ref)) <= 1e-5


@pytest.mark.gpu
def test_solver_backend_switch():
    """FK_DTI_Solver's default CUDA elongation stays within the fp64 gate of the literal torch restatement
    (elongation_backend='torch', the reference's own float32 CPU eigh)."""
    from tumorgrowthtoolkit_b200.FK_DTI import FK_DTI_Solver
    from tumorgrowthtoolkit_b200 import synthetic
    wm, gm = synthetic.brain_phantom((40, 44, 36))
    rng = np.random.default_rng(2)
    q, _ = np.linalg.qr(rng.normal(size=wm.shape + (3, 3)))
    lam = np.stack([0.3 + 0 * wm, 0.5 + 0 * wm, 1.7 * wm + 0.6], axis=-1) * 1e-3
    tensors = np.einsum("...ij,...j,...kj->...ik", q, lam, q) * ((wm + gm) > 0.1)[..., None, None]
    tensors = 0.5 * (tensors + np.swapaxes(tensors, -1, -2))
    base = dict(Dw=0.8, rho=0.3, diffusionEllipsoidScaling=1.8, diffusionTensors=tensors, NxT1_pct=0.5, NyT1_pct=0.5,
                NzT1_pct=0.5, resolution_factor=1.0, stopping_time=12, init_scale=1.0)
    a = FK_DTI_Solver(dict(base, elongation_backend='torch')).solve()
    b = FK_DTI_Solver(dict(base)).solve()
    assert a['success'] and b['success']
    assert np.max(np.abs(a['final_state'] - b['final_state'])) <= 1e-6
