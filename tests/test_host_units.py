"""Unit tests of the host-side contract arithmetic (CPU only): resample shapes, seed, step-count
law, crop boxes -- the few lines that decide which grid and how many steps the device loop sees."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from tumorgrowthtoolkit_b200 import _host


@settings(max_examples=60, deadline=None)
@given(shape=st.tuples(st.integers(2, 40), st.integers(2, 40), st.integers(2, 40)),
       factor=st.sampled_from([0.25, 0.5, 0.6, 0.65, 1.0, 1.5, 2.0, 1.7]))
def test_resample_output_shape_is_scipys(shape, factor):
    from scipy.ndimage import zoom
    assert _host._out_shape(shape, factor) == zoom(np.zeros(shape), factor, order=1).shape


@pytest.mark.parametrize("shape,rf,pct,scale,h", [((156, 156, 101), 0.65, (0.3, 0.7, 0.5), 1.0, (1, 1, 1)),
                                                  ((120, 120, 78), 0.5, (0.31, 0.02, 0.99), 1.0, (1, 1.2, 0.9)),
                                                  ((40, 44, 30), 1.0, (0.5, 0.5, 0.5), 0.1, (1, 1, 1)),
                                                  ((60, 60, 60), 0.6, (0.0, 0.999, 0.5), 3.0, (2, 1, 1))])
def test_windowed_seed_equals_the_dense_formula(shape, rf, pct, scale, h):
    """gauss_sol3d (FK/FK.py:92-105) evaluated on the whole grid, as the reference does, against
    the windowed evaluation: bitwise equal."""
    grid = _host.LowResGrid(shape, shape, rf, *h, pct)
    got = _host.seed_gaussian(grid, scale)
    (nx, ny, nz), (dx, dy, dz), (sx, sy, sz) = grid.shape, grid.h, grid.seed
    x, y, z = np.meshgrid(np.arange(nx) - sx, np.arange(ny) - sy, np.arange(nz) - sz, indexing="ij")
    g = 250 / np.power(4 * np.pi * 5.0, 3 / 2) * np.exp(
        -(np.power(x * dx / scale, 2) + np.power(y * dy / scale, 2) + np.power(z * dz / scale, 2)) / (4 * 5.0))
    g = np.where(g > 0.1, g, 0)
    g = np.where(g > 1, np.float64(1), g)
    assert np.array_equal(got, g)
    assert (got > 0).sum() > 0


def test_step_counts_printed_by_the_reference_notebooks():
    """Number of simulation timesteps: 300 (FK_exampleAtlas.ipynb:94), 560 (FK_2c.py:200 with the
    FK_2c_example.ipynb parameters; the notebook's saved '360' is stale)."""
    dx = 1.0 / 0.5
    nt = np.max([100 * 1.0 / np.power(dx, 2) * 8 + 100, 100 * 0.10 * 1.1])
    assert _host.steps_and_dt(nt, 100) == (300, 100 / nt)
    nt = 100 * np.max([0.9, 1.3]) / np.power(dx, 2) * 8 + 300
    assert _host.steps_and_dt(nt, 100)[0] == 560
    nt = 100 * 1.3 / 1.0 * 8 + 300
    assert _host.steps_and_dt(nt, 100)[0] == 1340
    # Nt is a float: dt = T/Nt but ceil(Nt) steps run (FK.py:185-187)
    steps, dt = _host.steps_and_dt(100.2, 2)
    assert steps == 101 and dt == 2 / 100.2


def test_record_schedule_and_boxes():
    assert _host.record_schedule(300, None) is None
    r = _host.record_schedule(300, 64)
    assert r[0] == 0 and r[-1] == 299 and len(r) == 64 and r.dtype.kind == "i"
    assert len(np.unique(_host.record_schedule(10, 200))) == 10          # more snapshots than steps
    m = np.zeros((10, 12, 9), dtype=bool)
    m[3:6, 1:4, 7:9] = True
    lo, hi = _host.bounding_box(m, 2, m.shape)
    assert list(lo) == [1, 0, 5] and list(hi) == [8, 6, 9]               # margin 2, clamped to the volume
    a = np.arange(m.size, dtype=float).reshape(m.shape)
    c = _host.crop(a, (lo, hi))
    assert c.shape == (7, 6, 4) and not c.flags["C_CONTIGUOUS"]
    back = _host.paste(m.shape, c, (lo, hi))
    assert np.array_equal(back[1:8, 0:6, 5:9], c) and back.sum() == c.sum()


def test_input_validation_raises_assertion_errors():
    ok = np.zeros((4, 4, 4))
    with pytest.raises(AssertionError):
        _host.check_tissue_inputs([[0]], ok, (0.5, 0.5, 0.5))
    with pytest.raises(AssertionError):
        _host.check_tissue_inputs(ok, np.zeros((4, 4)), (0.5, 0.5, 0.5))
    with pytest.raises(AssertionError):
        _host.check_tissue_inputs(ok, np.zeros((4, 4, 5)), (0.5, 0.5, 0.5))
    with pytest.raises(AssertionError):
        _host.check_tissue_inputs(ok, ok, (0.5, -0.1, 0.5))
    _host.check_tissue_inputs(ok, ok, (0.0, 1.0, 0.5))
