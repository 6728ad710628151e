"""The two exact per-warp shortcuts of the two-voxel FK / FK_2c kernels (tgtk_kernels_v3.cuh: warps without tumour skip the
Heaviside, warps that also lie outside the tissue skip the arithmetic) must not change a single bit: a box with background,
a closed shell, tissue and a seed whose front moves into tumour-free warps, stepped with TGTK_CLOSED_SKIP=1 (default) and 0."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def _phantom(shape, seed):
    from tumorgrowthtoolkit_b200 import synthetic
    wm, gm = synthetic.brain_phantom(shape)
    rng = np.random.default_rng(seed)
    X, Y, Z = shape
    x, y, z = np.meshgrid(np.arange(X) - X // 2, np.arange(Y) - Y // 2 + 3, np.arange(Z) - Z // 2, indexing="ij")
    u0 = np.clip(0.9 * np.exp(-(x * x + y * y + z * z) / 10.0), 0, 1)
    u0[u0 < 0.05] = 0.0
    u0 *= (wm + gm >= 0.1)
    return wm, gm, u0, rng


def _run(fn, *args, skip, **kw):
    os.environ["TGTK_CLOSED_SKIP"] = skip
    try:
        return fn(*args, **kw)
    finally:
        os.environ.pop("TGTK_CLOSED_SKIP", None)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("shape", [(40, 48, 70), (33, 41, 37)])
def test_fk2c_shortcuts_change_nothing(shape, precision):
    from tumorgrowthtoolkit_b200 import _loop
    wm, gm, u0, _ = _phantom(shape, 1)
    assert (wm + gm < 0.1).mean() > 0.3 and (u0 > 0).sum() > 20        # background and a seed are both there
    S0 = np.where(wm + gm >= 0.1, 1.0, 0.0)
    args = (u0, np.zeros_like(u0), S0, wm, gm, 0.1, 0.9, 0.9, 100.0, 1.3, 0.5, 0.35, 0.5, 2.0, 0.03, 1.0, 1.0, 1.0, 90)
    rec = np.arange(0, 90, 11)
    a = _run(_loop.fk2c_loop, *args, skip="1", precision=precision, record_steps=rec)
    b = _run(_loop.fk2c_loop, *args, skip="0", precision=precision, record_steps=rec)
    for k in "PNS":
        assert np.array_equal(a["state"][k], b["state"][k]), k
        for p, q in zip(a["snapshots"][k], b["snapshots"][k]):
            assert np.array_equal(p, q)
    assert a["volume"] == b["volume"]
    assert a["state"]["N"].max() > 0.1                                     # necrosis developed: the Heaviside mattered


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_fk_shortcut_changes_nothing(precision):
    from tumorgrowthtoolkit_b200 import _loop
    os.environ["TGTK_VARIANT"] = "4"
    try:
        wm, gm, u0, _ = _phantom((36, 50, 66), 2)
        args = (u0, wm, gm, 0.1, 1.0, 10.0, 0.2, 0.03, 1.0, 1.0, 1.0, 80)
        a = _run(_loop.fk_loop, *args, skip="1", precision=precision)
        b = _run(_loop.fk_loop, *args, skip="0", precision=precision)
    finally:
        os.environ.pop("TGTK_VARIANT", None)
    assert np.array_equal(a["state"], b["state"]) and a["volume"] == b["volume"]
