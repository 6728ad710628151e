"""Sparse launches (tgtk_kernels_sparse.cuh, include/tgtk_b200.h: tgtk_sim_sparse_info): warp tiles outside the tissue are never
loaded, computed or stored and the grid is a work list over the rest.  The results must be those of the dense launches bit
for bit: final state, snapshots, volume log, stop step."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def _problem(shape, seed_shift=3):
    from tumorgrowthtoolkit_b200 import synthetic
    wm, gm = synthetic.brain_phantom(shape)
    X, Y, Z = shape
    x, y, z = np.meshgrid(np.arange(X) - X // 2, np.arange(Y) - Y // 2 + seed_shift, np.arange(Z) - Z // 2, indexing="ij")
    u0 = np.clip(0.9 * np.exp(-(x * x + y * y + z * z) / 10.0), 0, 1)
    u0[u0 < 0.05] = 0.0
    S0 = np.where(wm + gm >= 0.1, 1.0, 0.0)
    return wm, gm, u0, S0


def _env(**kw):
    class _E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return _E()


def _args(wm, gm, u0, S0, nsteps, lambda_s=2.0):
    return (u0, np.zeros_like(u0), S0, wm, gm, 0.1, 0.9, 0.9, 100.0, 1.3, 0.5, 0.35, 0.5, lambda_s, 0.03, 1.0, 1.0, 1.0, nsteps)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("shape", [(40, 48, 70), (33, 41, 37), (24, 90, 130)])
def test_fk2c_sparse_equals_dense(shape, precision):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, S0 = _problem(shape)
    # the seed overlaps the background (tiles that are closed but hold tumour must stay live)
    assert ((u0 > 0) & (wm + gm < 0.1)).any() or shape[0] < 30
    rec = np.arange(0, 100, 13)
    with _env(TGTK_SPARSE=1):
        a = _loop.fk2c_loop(*_args(wm, gm, u0, S0, 100), precision=precision, record_steps=rec)
    with _env(TGTK_SPARSE=0):
        b = _loop.fk2c_loop(*_args(wm, gm, u0, S0, 100), precision=precision, record_steps=rec)
    assert a["sparse_items"] > 0 and b["sparse_items"] == 0
    assert 0 < a["sparse_live_fraction"] < 0.9 and b["sparse_live_fraction"] == 1.0
    for k in "PNS":
        assert np.array_equal(a["state"][k], b["state"][k]), k
        assert len(a["snapshots"][k]) == len(rec)
        for p, q in zip(a["snapshots"][k], b["snapshots"][k]):
            assert np.array_equal(p, q)
    assert abs(a["volume"] - b["volume"]) <= 1e-12 * abs(b["volume"])      # per-block partial sums: another grid, another order
    assert a["state"]["N"].max() > 0.1


@pytest.mark.parametrize("waves", [0.5, 1, 4])
def test_fk2c_sparse_item_sizes(waves):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, S0 = _problem((70, 64, 96))
    with _env(TGTK_SPARSE=1, TGTK_SP_WAVES=waves):
        a = _loop.fk2c_loop(*_args(wm, gm, u0, S0, 40))
    with _env(TGTK_SPARSE=0):
        b = _loop.fk2c_loop(*_args(wm, gm, u0, S0, 40))
    for k in "PNS":
        assert np.array_equal(a["state"][k], b["state"][k]), k


def test_fk2c_sparse_stop_criterion():
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, S0 = _problem((40, 48, 70))
    with _env(TGTK_SPARSE=0):
        full = _loop.fk2c_loop(*_args(wm, gm, u0, S0, 120, lambda_s=0.05))
    target = float(full["volume"]) * 0.8
    out = []
    for sparse in (1, 0):
        with _env(TGTK_SPARSE=sparse):
            out.append(_loop.fk2c_loop(*_args(wm, gm, u0, S0, 120, lambda_s=0.05), stopping_volume=target))
    a, b = out
    assert a["steps_run"] == b["steps_run"] and 10 < a["steps_run"] < 120 and a["stop"] == "volume"
    for k in "PNS":
        assert np.array_equal(a["state"][k], b["state"][k]), k


def test_fk2c_sparse_reset_and_new_state():
    """tgtk_sim_reset keeps the mask (same initial state); a new upload rebuilds it -- a seed placed where the first mask
    had dead tiles must grow."""
    from tumorgrowthtoolkit_b200 import _native as nat
    shape = (40, 48, 70)
    wm, gm, u0, S0 = _problem(shape)
    with _env(TGTK_SPARSE=1):
        p = nat.make_params(nat.MODEL_FK_2C, nat.F64, shape, (1.0, 1.0, 1.0), 0.03, rho=0.5, Dw=0.9, ratio=100.0, th_matter=0.1,
                            th_necro=0.9, Ds=1.3, lambda_np=0.35, sigma_np=0.5, lambda_s=2.0)
        sim = nat.Sim(p, 0)
    with _env(TGTK_SPARSE=0):
        ref = nat.Sim(p, 0)
    try:
        def load(s, P0):
            s.upload(nat.FIELD_P, P0); s.upload(nat.FIELD_N, np.zeros_like(P0)); s.upload(nat.FIELD_S, S0)
            s.upload(nat.FIELD_WM, wm); s.upload(nat.FIELD_GM, gm)
            s.prepare()
        for s in (sim, ref):
            load(s, u0)
            s.run(37); s.sync()
        assert sim.sparse_info()["items"] > 0 and ref.sparse_info()["items"] == 0
        first = {k: sim.download(f) for k, f in (("P", nat.FIELD_P), ("N", nat.FIELD_N), ("S", nat.FIELD_S))}
        for k, f in (("P", nat.FIELD_P), ("N", nat.FIELD_N), ("S", nat.FIELD_S)):
            assert np.array_equal(first[k], ref.download(f)), k
        for s in (sim, ref):
            s.reset()
            s.run(37); s.sync()
        for k, f in (("P", nat.FIELD_P), ("N", nat.FIELD_N), ("S", nat.FIELD_S)):
            assert np.array_equal(sim.download(f), first[k]), k
        # a different seed: far from the first one
        u1 = np.roll(u0, (9, -11, 14), axis=(0, 1, 2)) * (wm + gm >= 0.1)
        assert u1.max() > 0.3
        for s in (sim, ref):
            s.reset()
            s.upload(nat.FIELD_P, u1)
            s.run(45); s.sync()
        for k, f in (("P", nat.FIELD_P), ("N", nat.FIELD_N), ("S", nat.FIELD_S)):
            assert np.array_equal(sim.download(f), ref.download(f)), k
    finally:
        sim.close(); ref.close()


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("shape", [(40, 48, 70), (33, 41, 37)])
def test_fk_sparse_equals_dense(shape, precision):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, _ = _problem(shape)
    rec = np.arange(0, 90, 17)
    args = (u0, wm, gm, 0.1, 1.0, 10.0, 0.2, 0.03, 1.0, 1.0, 1.0, 90)
    with _env(TGTK_SPARSE=1, TGTK_VARIANT=4):
        a = _loop.fk_loop(*args, precision=precision, record_steps=rec)
    with _env(TGTK_SPARSE=0, TGTK_VARIANT=4):
        b = _loop.fk_loop(*args, precision=precision, record_steps=rec)
    assert a["sparse_items"] > 0 and b["sparse_items"] == 0
    assert np.array_equal(a["state"], b["state"])
    for p, q in zip(a["snapshots"], b["snapshots"]):
        assert np.array_equal(p, q)
    assert abs(a["volume"] - b["volume"]) <= 1e-12 * abs(b["volume"])
    assert (a["state"] > 0).sum() > 2 * (u0 > 0).sum()          # the tumour spread into tiles that held none


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("shape", [(40, 48, 70), (33, 41, 37)])
def test_dti_sparse_equals_dense(shape, precision):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, _ = _problem(shape)
    rng = np.random.default_rng(5)
    rgb = rng.uniform(0.05, 1.0, shape + (3,)) * (wm + gm >= 0.1)[..., None]
    u0 = u0 * (wm + gm >= 0.1)
    args = (u0, rgb, 1.0, 0.2, 0.03, 1.0, 1.0, 1.0, 90)
    out = []
    for sparse in (1, 0):
        with _env(TGTK_SPARSE=sparse, TGTK_VARIANT=4, TGTK_TB=0):
            out.append(_loop.dti_loop(*args, precision=precision, record_steps=np.arange(0, 90, 17)))
    a, b = out
    assert a["sparse_items"] > 0 and b["sparse_items"] == 0
    assert np.array_equal(a["state"], b["state"])
    for p, q in zip(a["snapshots"], b["snapshots"]):
        assert np.array_equal(p, q)
    assert abs(a["volume"] - b["volume"]) <= 1e-12 * abs(b["volume"])


@pytest.mark.parametrize("shape", [(9, 5, 7), (16, 3, 40), (8, 20, 33), (30, 9, 65)])
def test_sparse_on_boxes_narrower_than_a_tile(shape):
    """boxes smaller than one 8 x 32 tile in y or z (every halo wraps around inside the tile), forced sparse"""
    from tumorgrowthtoolkit_b200 import _loop
    rng = np.random.default_rng(7)
    X, Y, Z = shape
    wm = rng.uniform(0.2, 0.8, shape)
    gm = 1.0 - wm
    dead = np.zeros(shape, bool)
    dead[: X // 3] = True                                   # a third of the planes outside the tissue ...
    dead[:, :, Z // 2:] |= rng.uniform(size=(X, Y, Z - Z // 2)) < 0.3     # ... and scattered background elsewhere
    wm[dead] = 0.0
    gm[dead] = 0.0
    u0 = np.zeros(shape)
    u0[X // 2, Y // 2, Z // 4] = 0.8
    u0[X // 2, Y // 2, (Z // 4 + 1) % Z] = 0.4
    S0 = np.where(wm + gm >= 0.1, 1.0, 0.0)
    out = {}
    for sparse in (1, 0):
        with _env(TGTK_SPARSE=sparse, TGTK_VARIANT=4):
            out[sparse] = (_loop.fk2c_loop(*_args(wm, gm, u0, S0, 50)), _loop.fk_loop(u0, wm, gm, 0.1, 1.0, 10.0, 0.2, 0.03, 1.0, 1.0, 1.0, 50))
    assert out[1][0]["sparse_items"] > 0 and out[1][1]["sparse_items"] > 0
    for k in "PNS":
        assert np.array_equal(out[1][0]["state"][k], out[0][0]["state"][k]), k
    assert np.array_equal(out[1][1]["state"], out[0][1]["state"])
    assert (out[1][1]["state"] > 0).sum() > 10
