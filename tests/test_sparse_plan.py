"""Host logic of the sparse launches: the work-list planner (tgtk_sparse_plan_host, include/tgtk_b200.h) on synthetic tile
masks -- no GPU involved.  What the step kernels rely on: every live tile of the planes to compute lies in exactly one item,
items stay inside those planes and never start or end on a dead plane, the list fits one round of resident blocks when a
slot's share is short, and in a slab the items that touch a ghost plane come last."""
import os

import numpy as np
import pytest

from tumorgrowthtoolkit_b200 import _native as nat


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as ge
    ge.build()


def _brain_mask(X, nzt, nyt, seed, hole=True):
    """an ellipsoid of live tiles with ragged warp bytes, optionally a dead cavity (ventricle) inside"""
    rng = np.random.default_rng(seed)
    x, z, y = np.meshgrid(np.linspace(-1, 1, X), np.linspace(-1, 1, nzt), np.linspace(-1, 1, nyt), indexing="ij")
    r = x * x / 0.8 + z * z / 0.9 + y * y / 0.7
    live = r < 1.0
    if hole:
        live &= ~((np.abs(x) < 0.12) & (np.abs(y) < 0.3) & (np.abs(z) < 0.3))
    words = np.zeros((X, nzt, nyt), dtype=np.uint32)
    for w in range(4):
        on = live & (rng.uniform(size=live.shape) < 0.85)
        words |= on.astype(np.uint32) << np.uint32(8 * w)
    words[live & (words == 0)] = 1
    return words


def _check_cover(mask, items, x_lo, x_hi):
    X, nzt, nyt = mask.shape
    cover = np.zeros((X, nzt, nyt), dtype=np.int32)
    for yt, zt, x0, x1 in items:
        assert 0 <= yt < nyt and 0 <= zt < nzt and x_lo <= x0 < x1 <= x_hi
        assert mask[x0, zt, yt] != 0 and mask[x1 - 1, zt, yt] != 0, "an item starts or ends on a dead plane"
        cover[x0:x1, zt, yt] += 1
    live = mask != 0
    live[:x_lo] = False
    live[x_hi:] = False
    assert (cover[live] == 1).all(), "a live tile is missing from the list or listed twice"
    assert cover.max() <= 1


@pytest.mark.parametrize("shape,slots", [((145, 5, 23), 592), ((145, 5, 23), 1184), ((36, 9, 45), 592), ((288, 9, 45), 1184),
                                         ((20, 2, 3), 592), ((9, 1, 1), 4)])
def test_plan_covers_every_live_tile_once(shape, slots):
    mask = _brain_mask(*shape, seed=sum(shape))
    if not mask.any():                       # tiny boxes: the ellipsoid misses the grid points -- everything live instead
        mask[1:-1] = 0x01000001
    items = nat.sparse_plan(mask, slots=slots)
    assert len(items) > 0
    _check_cover(mask, items, 0, shape[0])


def test_plan_short_shares_fit_one_round():
    mask = _brain_mask(145, 5, 23, seed=1)
    for slots in (592, 1184):
        items = nat.sparse_plan(mask, slots=slots)
        assert len(items) <= slots
        planes = items[:, 3] - items[:, 2]
        assert planes.min() >= 1 and planes.max() <= 3 * np.median(planes) + 8      # no straggler item


def test_plan_long_shares_balance():
    """a big box: several rounds of items, none longer than a few rounds' worth"""
    mask = _brain_mask(288, 9, 45, seed=2)
    items = nat.sparse_plan(mask, slots=1184)
    _check_cover(mask, items, 0, 288)
    planes = (items[:, 3] - items[:, 2]).astype(float)
    share = planes.sum() / 1184
    assert planes.max() <= 1.5 * share + 8


def test_plan_slab_range_and_order():
    mask = _brain_mask(38, 9, 44, seed=3, hole=False)
    mask[:] = mask[19]                       # a central slab: the tissue runs through every plane, ghosts included
    mask[10:22, :, :20] = 0                  # ... except for a cavity, so that some columns have interior-only items
    items = nat.sparse_plan(mask, x_lo=1, x_hi=37, slots=592, slab=True)
    _check_cover(mask, items, 1, 37)
    boundary = (items[:, 2] == 1) | (items[:, 3] == 37)
    assert boundary.any() and not boundary.all()
    first_boundary = int(np.argmax(boundary))
    assert boundary[first_boundary:].all(), "an interior item is listed after an item that waits for a neighbour"


def test_plan_gaps_split_columns():
    """two lobes in one column, far apart: two items, nothing scheduled for the gap"""
    mask = np.zeros((60, 1, 1), dtype=np.uint32)
    mask[5:15] = 0x01010101
    mask[40:52] = 0x00000100
    items = nat.sparse_plan(mask, slots=8)
    got = sorted((int(a), int(b)) for _, _, a, b in items)
    assert got[0][0] == 5 and got[-1][1] == 52
    assert all(not (a < 30 < b) for a, b in got)
    _check_cover(mask, items, 0, 60)


def test_plan_all_dead_and_waves_override():
    mask = np.zeros((16, 2, 2), dtype=np.uint32)
    assert len(nat.sparse_plan(mask, slots=16)) == 0
    mask = _brain_mask(145, 5, 23, seed=4)
    os.environ["TGTK_SP_WAVES"] = "3"
    try:
        many = nat.sparse_plan(mask, slots=592)
    finally:
        os.environ.pop("TGTK_SP_WAVES", None)
    one = nat.sparse_plan(mask, slots=592)
    assert len(many) > len(one)
    _check_cover(mask, many, 0, 145)


def test_plan_rejects_bad_arguments():
    mask = np.ones((8, 1, 1), dtype=np.uint32)
    for kw in (dict(x_lo=5, x_hi=5), dict(x_lo=0, x_hi=9), dict(slots=0)):
        with pytest.raises(nat.NativeError):
            nat.sparse_plan(mask, **kw)


def test_plan_items_are_ordered_longest_first():
    mask = _brain_mask(145, 5, 23, seed=5)
    items = nat.sparse_plan(mask, slots=592)
    planes = items[:, 3] - items[:, 2]
    # the list is sorted by modelled cost, which grows with the plane count: no late item is much longer than an early one
    assert planes[: len(planes) // 4].mean() >= planes[-len(planes) // 4:].mean()
