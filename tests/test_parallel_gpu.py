"""Row-stripe sharding end to end on the GPU (dask_geomodeling_b200/parallel.py): two
ranks -- NCCL with one GPU each when the box has two GPUs, otherwise gloo with both ranks
on cuda:0 -- must reproduce the single-process oracle result for a fused chain, the
stencils (halo exchange) and zonal statistics (all-reduce of partials, segment routing for
percentiles)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, n_gpus, fn):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = rank if n_gpus >= world else 0
    os.environ["GM_DEVICE"] = str(device)
    torch.cuda.set_device(device)
    if n_gpus >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def spawn(fn, world=2):
    import torch
    import torch.multiprocessing as mp

    mp.spawn(_run, args=(world, _free_port(), torch.cuda.device_count(), fn), nprocs=world, join=True)


def _chain_worker(rank, world):
    from dask_geomodeling_b200 import parallel, workloads
    from oracle import workloads as oracle_workloads

    size = 515  # odd: stripes of different heights
    ints, floats = workloads.cfg2_arrays(size, chunk=128)
    isdata, step = workloads.cfg2_views(ints, floats)
    (e_isdata, _), (e_step, _) = oracle_workloads.cfg2(ints, floats, workloads.CFG2_PAIRS)
    request = workloads.request(size, size)
    local, (r0, r1) = parallel.get_data_striped(step, **request)
    np.testing.assert_array_equal(local["values"], e_step[:, r0:r1])
    full, rows = parallel.get_data_striped(isdata, gather=True, **request)
    assert rows == (0, size)
    np.testing.assert_array_equal(full["values"], e_isdata)


def test_fused_chain_in_stripes():
    spawn(_chain_worker)


def _tiler_worker(rank, world):
    from dask_geomodeling_b200 import parallel, workloads
    from oracle import workloads as oracle_workloads

    size = 300
    ints, floats = workloads.cfg2_arrays(size, chunk=128)
    isdata, step = workloads.cfg2_views(ints, floats)
    (e_isdata, _), (e_step, _) = oracle_workloads.cfg2(ints, floats, workloads.CFG2_PAIRS)
    request = workloads.request(size, size)
    got = parallel.get_data_tiled(step, [128, 77], **request)       # 3 x 4 tiles over the ranks
    np.testing.assert_array_equal(got["values"], e_step)
    got = parallel.get_data_tiled(isdata, 64, **request)
    np.testing.assert_array_equal(got["values"], e_isdata)


def test_raster_tiler_over_ranks():
    # RasterTiler as the multi-GPU scheduler, against the oracle
    spawn(_tiler_worker)


def _stencil_worker(rank, world):
    import torch

    from dask_geomodeling_b200 import parallel, raster
    from oracle import raster as R

    rng = np.random.default_rng(3)
    h, w = 97, 131
    y, x = np.mgrid[0:h, 0:w]
    dem = (50 * np.sin(x / 17.0) + 30 * np.cos(y / 11.0) + rng.normal(0, 1, (2, h, w)) + 100).astype("f4")
    nodata = float(np.finfo("f4").max)
    dem[rng.random(dem.shape) < 0.03] = nodata
    r0, r1 = parallel.stripe_rows(h, world)[rank]
    local = torch.from_numpy(dem[:, r0:r1].copy()).cuda()

    def whole(halo):  # what one process sees: the raster padded with no data
        return np.pad(dem, ((0, 0), (halo, halo), (halo, halo)), constant_values=nodata)

    got = parallel.stencil_striped(raster.MovingMax.process, local, nodata, 5, 5, 11)
    expected, _ = R.moving_max(whole(5), nodata, 11)
    np.testing.assert_array_equal(np.asarray(got["values"]), expected[:, r0:r1])

    # Smooth: the stripe needs the full Gaussian radius (7 rows), not only the cropped margin (5)
    lw = parallel.smooth_halo(5.0)
    assert lw == 7
    kwargs = dict(smooth_mode="exact", fill=0, size=[5.0, 5.0], margin=(lw, 5))
    from dask_geomodeling_b200 import _native

    with _native.smooth_arithmetic("exact"):     # bit for bit against SciPy (the default is FMA)
        got = parallel.stencil_striped(raster.Smooth.process, local, nodata, lw, 5, kwargs)
    expected, _ = R.smooth(whole(5), nodata, (5.0, 5.0), 0, "exact")
    np.testing.assert_array_equal(np.asarray(got["values"]), expected[:, r0:r1])

    kwargs = dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0)
    got = parallel.stencil_striped(raster.HillShade.process, local, nodata, 1, 1, kwargs)
    expected, _ = R.hillshade(whole(1), nodata, (1.0, 1.0), 45.0, 315.0, 0)
    delta = np.abs(np.asarray(got["values"]).astype(int) - expected[:, r0:r1].astype(int))
    assert delta.max() <= 1 and (delta > 0).mean() <= 1e-3

    # a stripe STORED with its halo (one band): the exchange overlapped with the interior rows
    # (three row windows written into one output) equals the one-call result bit for bit, and
    # the oracle
    band = local[:1].contiguous()
    cases = [(raster.MovingMax.process, 5, 5, (11,), lambda: R.moving_max(whole(5)[:1], nodata, 11)[0]),
             (raster.HillShade.process, 1, 1, (dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0),), None),
             (raster.Smooth.process, lw, 5, (dict(smooth_mode="exact", fill=0, size=[5.0, 5.0], margin=(lw, 5)),),
              lambda: R.smooth(whole(5)[:1], nodata, (5.0, 5.0), 0, "exact")[0])]
    for process, halo_rows, halo_cols, extra, oracle in cases:
        pitch = 4 if process is raster.MovingMax.process else 1
        stored = parallel.pad_columns(parallel.exchange_halo(band, halo_rows, nodata), halo_cols, nodata, pitch)
        if process is raster.MovingMax.process:
            extra = extra + (stored.shape[2] - (w + 2 * halo_cols),)
        with _native.smooth_arithmetic("exact"):
            plain = np.asarray(parallel.stencil_haloed(process, stored, nodata, halo_rows, halo_cols, *extra,
                                                       overlap=False)["values"])
            stored[:, :halo_rows] = 0.0          # stale halo rows: the exchange must refresh them
            stored[:, -halo_rows:] = 0.0
            if rank == 0:
                stored[:, :halo_rows] = nodata
            if rank == world - 1:
                stored[:, -halo_rows:] = nodata
            over = parallel.stencil_haloed(process, stored, nodata, halo_rows, halo_cols, *extra, overlap=True)
        assert over["values"].shape == (1, r1 - r0, w)
        np.testing.assert_array_equal(np.asarray(over["values"]), plain)
        if oracle is not None:
            np.testing.assert_array_equal(plain, oracle()[:, r0:r1])


def test_stencils_with_halo_exchange():
    spawn(_stencil_worker)


def _zonal_worker(rank, world, on_device=False):
    import torch

    from dask_geomodeling_b200 import geometry, parallel, utils, workloads

    rng = np.random.default_rng(21)
    h, w = 203, 180
    frame = rng.uniform(0, 100, (1, h, w)).astype("f4")
    nodata = float(np.finfo("f4").max)
    frame[rng.random(frame.shape) < 0.1] = nodata
    polys = []
    for _ in range(37):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        k = int(rng.integers(3, 11))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(3, 60)   # up to 120 rows: crosses one stripe boundary, or two with three ranks
        ring = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
        polys.append(utils.Polygon(np.round(ring, 3) + 0.0137))
    polys.append(utils.Polygon([(5.01, 5.01), (5.02, 5.01), (5.02, 5.02)]))  # covers no cell centre
    bbox = (0, 0, w, h)
    r0, r1 = parallel.stripe_rows(h, world)[rank]
    local = np.ascontiguousarray(frame[:, r0:r1])
    if on_device:   # stripes resident in HBM: partials and boundary rows never touch the host
        local = torch.from_numpy(local).cuda()
    # the oracle: GDAL-rule labels per polygon + scipy labelled statistics / measurements.percentile
    from oracle import polyfill
    from oracle import raster as R

    rings = [[np.asarray(p.exterior.coords)] for p in polys]
    label_sets = []
    for i, ring in enumerate(rings):
        labels = polyfill.burn_index([ring], bbox, h, w)
        label_sets.append((np.where(labels == 0, i, np.iinfo(np.int32).max).astype(np.int32), [i]))
    for stat, q in (("mean", None), ("max", None), ("count", None), ("sum", None), ("min", None),
                    ("median", None), ("percentile", 90.0), ("percentile", 12.5)):
        expected, expected_no_cells = R.zonal_from_labels(frame[0], nodata, label_sets, len(polys), stat, q)
        got, no_cells = parallel.zonal_striped(polys, local, nodata, bbox, h, (r0, r1), stat, q)
        assert got.dtype == np.float32 and sorted(no_cells) == sorted(expected_no_cells)
        if stat in ("sum", "mean"):
            np.testing.assert_allclose(got, expected, rtol=1e-6, equal_nan=True)
        else:
            np.testing.assert_array_equal(got, expected)


def _zonal_worker_device(rank, world):
    _zonal_worker(rank, world, on_device=True)


def test_zonal_statistics_in_stripes():
    spawn(_zonal_worker)


def test_zonal_statistics_in_stripes_resident():
    spawn(_zonal_worker_device)


def test_zonal_statistics_in_three_stripes():
    # polygons that reach beyond the neighbouring stripe take the value-exchange path
    spawn(_zonal_worker, world=3)
