"""World-size-2 gloo tests of the multi-GPU plumbing (dask_geomodeling_b200/parallel.py):
stripe geometry, stencil halo exchange, all-reduce of zonal partials and the routing of
per-polygon value segments to their owner ranks.  No GPU and no kernel call: the
collectives run on CPU tensors, the expected values come from NumPy on the whole raster."""
import os
import socket

import numpy as np
import pytest

from dask_geomodeling_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, args):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def spawn(fn, *args, world=2):
    import torch.multiprocessing as mp

    mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=True)


# ---- pure geometry --------------------------------------------------------------------


@pytest.mark.parametrize("height,world", [(10, 2), (11, 4), (3, 8), (40000, 8)])
def test_stripe_rows_cover_the_raster(height, world):
    bounds = parallel.stripe_rows(height, world)
    assert bounds[0][0] == 0 and bounds[-1][1] == height
    assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
    sizes = [b - a for a, b in bounds]
    assert max(sizes) - min(sizes) <= 1


def test_stripe_request_geometry():
    request = dict(mode="vals", bbox=(100.0, 200.0, 140.0, 230.0), width=40, height=30, projection="EPSG:28992")
    subs = [parallel.stripe_request(request, r, 4) for r in range(4)]
    assert subs[0][0]["bbox"][3] == 230.0 and subs[-1][0]["bbox"][1] == 200.0  # row 0 is north
    for (a, ra), (b, rb) in zip(subs, subs[1:]):
        assert a["bbox"][1] == b["bbox"][3] and ra[1] == rb[0]
    for sub, (r0, r1) in subs:
        assert sub["height"] == r1 - r0 and sub["width"] == 40
        assert (sub["bbox"][3] - sub["bbox"][1]) / sub["height"] == 1.0  # pixel size unchanged


# ---- halo exchange ------------------------------------------------------------------------


def _halo_worker(rank, world, halo):
    import torch

    full = np.arange(3 * 23 * 7, dtype=np.float32).reshape(3, 23, 7)
    fill = -1.0
    r0, r1 = parallel.stripe_rows(23, world)[rank]
    got = parallel.exchange_halo(torch.from_numpy(full[:, r0:r1].copy()), halo, fill).numpy()
    padded = np.pad(full, ((0, 0), (halo, halo), (0, 0)), constant_values=fill)
    np.testing.assert_array_equal(got, padded[:, r0:r1 + 2 * halo])
    cols = parallel.pad_columns(torch.from_numpy(got), 2, fill).numpy()
    np.testing.assert_array_equal(cols[:, :, 2:-2], got)
    assert (cols[:, :, :2] == fill).all() and (cols[:, :, -2:] == fill).all()


    # the in-place variant on a stripe stored with its halo gives the same array
    stored = parallel.pad_columns(parallel.exchange_halo(torch.from_numpy(full[:, r0:r1].copy()), halo, fill), 2, fill)
    stored[:, :halo] = fill
    stored[:, -halo:] = fill
    parallel.refresh_halo(stored, halo, 2)
    np.testing.assert_array_equal(stored.numpy(), cols)


@pytest.mark.parametrize("halo", [1, 5, 7])
def test_exchange_halo_two_ranks(halo):
    spawn(_halo_worker, halo)


def test_exchange_halo_three_ranks():
    spawn(_halo_worker, 3, world=3)


def _box_max(data, size):
    """A window-invariant stencil with the blocks' calling convention: the input carries a margin
    of size // 2 cells that the output drops."""
    from scipy import ndimage

    r = size // 2
    out = ndimage.maximum_filter(np.asarray(data["values"]), size=(1, size, size))[:, r:-r, r:-r]
    return {"values": out, "no_data_value": data["no_data_value"]}


def _overlap_worker(rank, world, halo):
    import torch

    rng = np.random.default_rng(9)
    full = rng.normal(size=(1, 61, 23)).astype("f4")
    fill = -1e30
    r0, r1 = parallel.stripe_rows(full.shape[1], world)[rank]
    stored = parallel.pad_columns(parallel.exchange_halo(torch.from_numpy(full[:, r0:r1].copy()), halo, fill), halo, fill)
    expected = _box_max({"values": np.pad(full, ((0, 0), (halo, halo), (halo, halo)), constant_values=fill),
                         "no_data_value": fill}, 2 * halo + 1)["values"][:, r0:r1]
    for overlap in (False, True):
        stored[:, :halo] = 7.0       # stale halo rows: the exchange must refresh them
        stored[:, -halo:] = 7.0
        if rank == 0:
            stored[:, :halo] = fill
        if rank == world - 1:
            stored[:, -halo:] = fill
        got = parallel.stencil_haloed(_box_max, stored, fill, halo, halo, 2 * halo + 1, overlap=overlap)
        np.testing.assert_array_equal(np.asarray(got["values"]), expected)


@pytest.mark.parametrize("halo,world", [(1, 2), (5, 2), (3, 3)])
def test_stencil_on_a_stored_stripe_with_overlapped_exchange(halo, world):
    spawn(_overlap_worker, halo, world=world)


def _thin_stripe_worker(rank, world):
    import torch

    # 7 rows over 3 ranks: stripes of 3, 2, 2 rows; a halo of 3 rows fits only the first one.
    # EVERY rank must raise (a rank raising alone would leave the others in the send/recv)
    r0, r1 = parallel.stripe_rows(7, world)[rank]
    local = torch.zeros((1, r1 - r0, 4))
    with pytest.raises(ValueError, match="thinner than the halo"):
        parallel.exchange_halo(local, 3, -1.0)
    stored = torch.zeros((1, r1 - r0 + 6, 4))
    with pytest.raises(ValueError, match="thinner than the halo"):
        parallel.refresh_halo(stored, 3)
    got = parallel.exchange_halo(local, 2, -1.0)          # 2 rows fit every stripe
    assert got.shape == (1, r1 - r0 + 4, 4)


def test_thin_stripes_fail_on_every_rank():
    spawn(_thin_stripe_worker, world=3)


class _HalfEmptyView(object):
    """get_data answers None for requests in the southern half (y < 5), values elsewhere."""

    def get_data(self, **request):
        x1, y1, x2, y2 = request["bbox"]
        if y2 <= 5:
            return None
        return {"values": np.full((2, request["height"], request["width"]), 3, dtype="u1"), "no_data_value": 255}


def _gather_worker(rank, world):
    request = dict(mode="vals", bbox=(0, 0, 4, 10), width=4, height=10, projection="EPSG:28992")
    full, rows = parallel.get_data_striped(_HalfEmptyView(), gather=True, **request)
    assert rows == (0, 10) and full["values"].shape == (2, 10, 4)      # the height is kept
    assert (full["values"][:, :5] == 3).all() and (full["values"][:, 5:] == 255).all()


def test_gather_fills_stripes_without_data():
    spawn(_gather_worker)


class _CoordinateView(object):
    """A view whose cell value is a function of the cell's position: 100 * row + column of a
    40 x 30 grid with unit cells (no data outside x < 25)."""

    dtype = np.dtype("i4")
    fillvalue = np.iinfo("i4").max

    def get_data(self, **request):
        if request["mode"] != "vals":
            return {request["mode"]: ["whole"]}
        x1, y1, x2, y2 = request["bbox"]
        cols = np.arange(int(x1), int(x2))
        rows = np.arange(30 - int(y2), 30 - int(y1))
        values = (100 * rows[:, None] + cols[None, :]).astype("i4")
        values[:, cols >= 25] = self.fillvalue
        return {"values": values[np.newaxis], "no_data_value": self.fillvalue}


def _tiler_worker(rank, world):
    from dask_geomodeling_b200.raster.base import RasterBlock

    view = type("V", (_CoordinateView, RasterBlock), {"__init__": lambda self: None, "token": "v"})()
    request = dict(mode="vals", bbox=(2, 3, 40, 28), width=38, height=25, projection="EPSG:28992")
    expected = _CoordinateView().get_data(**request)
    for tile_size in (7, [16, 9], 100):
        got = parallel.get_data_tiled(view, tile_size, **request)
        np.testing.assert_array_equal(got["values"], expected["values"])
        assert got["no_data_value"] == expected["no_data_value"]
    assert parallel.get_data_tiled(view, 8, **dict(request, mode="meta")) == {"meta": ["whole"]}


def test_raster_tiler_over_ranks():
    spawn(_tiler_worker)
    spawn(_tiler_worker, world=3)


# ---- zonal partials ---------------------------------------------------------------------------


def _labelled_raster():
    rng = np.random.default_rng(5)
    values = rng.uniform(0, 100, (31, 17)).astype(np.float32)
    labels = rng.integers(0, 6, (31, 17))          # label 5 never has active cells (see below)
    active = rng.random((31, 17)) > 0.2
    active[labels == 5] = False
    return values, labels, active, 7                # label 6 covers no cell at all


def _partials_of(values, labels, active, n):
    out = np.zeros(n, dtype=parallel.PARTIAL_DTYPE)
    out["vmin"], out["vmax"] = np.finfo(np.float64).max, -np.finfo(np.float64).max
    covered = np.zeros(n, dtype=np.int64)
    for p in range(n):
        covered[p] = (labels == p).sum()
        v = values[(labels == p) & active].astype(np.float64)
        if v.size:
            out[p] = (v.size, v.sum(), v.min(), v.max())
    return out, covered


def _partials_worker(rank, world):
    values, labels, active, n = _labelled_raster()
    r0, r1 = parallel.stripe_rows(values.shape[0], world)[rank]
    partial, covered = _partials_of(values[r0:r1], labels[r0:r1], active[r0:r1], n)
    reduced, covered = parallel.allreduce_partials(partial, covered)
    whole, whole_covered = _partials_of(values, labels, active, n)
    np.testing.assert_array_equal(reduced["count"], whole["count"])
    np.testing.assert_array_equal(covered, whole_covered)
    np.testing.assert_array_equal(reduced["vmin"], whole["vmin"])
    np.testing.assert_array_equal(reduced["vmax"], whole["vmax"])
    np.testing.assert_allclose(reduced["sum"], whole["sum"], rtol=1e-14)
    for stat in ("count", "sum", "mean", "min", "max"):
        got = parallel.finalize_partials(reduced, stat)
        assert got.dtype == np.float32
        for p in range(n):
            v = values[(labels == p) & active].astype(np.float64)
            if v.size == 0:
                assert np.isnan(got[p])
                continue
            expected = {"count": v.size, "sum": v.sum(), "mean": v.sum() / v.size, "min": v.min(), "max": v.max()}[stat]
            np.testing.assert_allclose(got[p], np.float32(expected), rtol=1e-6)


def test_allreduce_partials_two_ranks():
    spawn(_partials_worker)


# ---- segment routing for order statistics ------------------------------------------------------


def _segments_worker(rank, world):
    values, labels, active, n = _labelled_raster()
    r0, r1 = parallel.stripe_rows(values.shape[0], world)[rank]
    counts = np.array([((labels[r0:r1] == p) & active[r0:r1]).sum() for p in range(n)], dtype=np.int64)
    packed = np.concatenate([values[r0:r1][(labels[r0:r1] == p) & active[r0:r1]] for p in range(n)])
    owned, offsets, merged = parallel.exchange_segments(counts, packed)
    np.testing.assert_array_equal(owned, np.arange(rank, n, world))
    assert merged.dtype == values.dtype
    for k, p in enumerate(owned):
        segment = merged[offsets[k]:offsets[k + 1]]
        np.testing.assert_array_equal(np.sort(segment), np.sort(values[(labels == p) & active]))
    # the owner's percentile over the merged segment is the whole-raster percentile
    # (measurements.py:132-137 restated with NumPy; the product uses gm_segment_order_stat)
    mine = np.full(len(owned), np.nan, dtype=np.float32)
    for k in range(len(owned)):
        d = np.sort(merged[offsets[k]:offsets[k + 1]])
        if d.size:
            frac = (d.size - 1) * 0.9
            lo, hi = int(np.floor(frac)), int(np.ceil(frac))
            mine[k] = np.float64(d[lo]) + (frac % 1) * np.float64(d[hi] - d[lo])
    everyone = parallel._gather_owned(mine, owned, n)
    for p in range(n):
        d = np.sort(values[(labels == p) & active])
        if d.size == 0:
            assert np.isnan(everyone[p])
        else:
            frac = (d.size - 1) * 0.9
            lo, hi = int(np.floor(frac)), int(np.ceil(frac))
            assert everyone[p] == np.float32(np.float64(d[lo]) + (frac % 1) * np.float64(d[hi] - d[lo]))


def test_exchange_segments_two_ranks():
    spawn(_segments_worker)


def test_exchange_segments_three_ranks():
    spawn(_segments_worker, world=3)


def test_single_process_paths():
    # without an initialised process group everything degenerates to one stripe
    counts = np.array([2, 0, 1], dtype=np.int64)
    values = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    owned, offsets, merged = parallel.exchange_segments(counts, values)
    np.testing.assert_array_equal(owned, [0, 1, 2])
    np.testing.assert_array_equal(offsets, [0, 2, 2, 3])
    np.testing.assert_array_equal(merged, values)
    sub, rows = parallel.stripe_request(dict(bbox=(0, 0, 4, 6), height=6, width=4), 0, 1)
    assert rows == (0, 6) and sub["bbox"] == (0, 0, 4, 6)
