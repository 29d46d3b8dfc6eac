"""GPU scanline rasteriser and zonal statistics against the oracle (GDAL fill
restatement + scipy.ndimage / measurements.percentile on label rasters)."""
import numpy as np
import pandas as pd
import pytest

from dask_geomodeling_b200 import geometry, raster, utils, workloads
from oracle import polyfill
from oracle import raster as R

pytestmark = pytest.mark.gpu


def random_polygons(n, size, seed, concave=False):
    """n-gons scattered over [0, size]^2, vertices never on k + 0.5."""
    rng = np.random.default_rng(seed)
    polys = []
    for _ in range(n):
        cx, cy = rng.uniform(0, size, 2)
        k = int(rng.integers(3, 13))
        angles = np.sort(rng.uniform(0, 2 * np.pi, k))
        radius = rng.uniform(2, size / 6) * (rng.uniform(0.3, 1.0, k) if concave else 1.0)
        ring = np.stack([cx + radius * np.cos(angles), cy + radius * np.sin(angles)], axis=1)
        polys.append([np.round(ring, 3) + 0.0137])
    return polys


def to_geometries(polys):
    return [utils.Polygon(rings[0], rings[1:]) for rings in polys]


@pytest.mark.parametrize("concave", [False, True])
@pytest.mark.parametrize("shape", [(64, 80), (301, 257)])
def test_rasterize_matches_gdal_rule(shape, concave):
    h, w = shape
    polys = random_polygons(60, max(h, w), seed=h + concave, concave=concave)
    bbox = (0, 0, w, h)
    ids = np.arange(100, 100 + len(polys))
    expected, nodata = polyfill.rasterize(polys, bbox, h, w, values=ids)
    got = utils.rasterize_geoseries(to_geometries(polys), bbox, workloads.PROJECTION, h, w,
                                    values=pd.Series(ids))
    assert got["values"].dtype == np.int32 and got["no_data_value"] == nodata
    np.testing.assert_array_equal(got["values"], expected)
    got = utils.rasterize_geoseries(to_geometries(polys), bbox, workloads.PROJECTION, h, w)
    np.testing.assert_array_equal(got["values"], expected != nodata)
    assert got["values"].dtype == bool and got["no_data_value"] is None
    floats = ids / 7.0
    expected_f, nodata_f = polyfill.rasterize(polys, bbox, h, w, values=floats)
    got = utils.rasterize_geoseries(to_geometries(polys), bbox, workloads.PROJECTION, h, w,
                                    values=pd.Series(floats))
    assert got["values"].dtype == np.float64
    np.testing.assert_array_equal(got["values"], expected_f)


def test_rasterize_holes_multipolygon_and_many_vertices():
    shell = [(1, 1), (59, 1), (59, 59), (1, 59)]
    hole = [(10.2, 10.2), (30.7, 10.2), (30.7, 40.1), (10.2, 40.1)]
    t = np.linspace(0, 2 * np.pi, 300, endpoint=False)
    blob = np.round(np.stack([30 + (20 + 5 * np.sin(7 * t)) * np.cos(t),
                              30 + (20 + 5 * np.sin(7 * t)) * np.sin(t)], axis=1), 3) + 0.0137
    multi = utils.MultiPolygon([utils.Polygon([(2, 2), (8.3, 2), (8.3, 8.3)]),
                                utils.Polygon([(40, 40), (55.5, 40), (55.5, 55.5), (40, 55.5)])])
    geoms = [utils.Polygon(shell, [hole]), utils.Polygon(blob), multi]
    rings = [[shell, hole], [blob], [[(2, 2), (8.3, 2), (8.3, 8.3)], [(40, 40), (55.5, 40), (55.5, 55.5), (40, 55.5)]]]
    expected, nodata = polyfill.rasterize(rings, (0, 0, 60, 60), 60, 60, values=[1, 2, 3])
    got = utils.rasterize_geoseries(geoms, (0, 0, 60, 60), workloads.PROJECTION, 60, 60,
                                    values=pd.Series([1, 2, 3]))
    np.testing.assert_array_equal(got["values"], expected)


def oracle_zonal(frame, nodata, polys, bbox, statistic, q=None, thresholds=None):
    h, w = frame.shape
    label_sets = []
    for i, rings in enumerate(polys):  # one bucket per polygon: overlap-proof
        labels = polyfill.burn_index([rings], bbox, h, w)
        labels = np.where(labels == 0, i, np.iinfo(np.int32).max).astype(np.int32)
        label_sets.append((labels, [i]))
    return R.zonal_from_labels(frame, nodata, label_sets, len(polys), statistic, q, thresholds)


@pytest.mark.parametrize("dtype", ["f4", "u1", "i2", "i4", "f8"])
@pytest.mark.parametrize("statistic", ["sum", "count", "min", "max", "mean", "median", "p90", "p12.5"])
def test_zonal_stats(dtype, statistic):
    h, w = 120, 150
    rng = np.random.default_rng(11)
    nodata = R.dtype_max(dtype)
    frame = (rng.uniform(0, 100, (h, w)) if dtype in ("f4", "f8") else rng.integers(0, 100, (h, w))).astype(dtype)
    frame[rng.random((h, w)) < 0.1] = nodata
    polys = random_polygons(40, 150, seed=3, concave=True)
    polys.append([[(200.0, 200.0), (210.0, 200.0), (210.0, 210.0)]])      # outside the raster
    polys.append([[(5.01, 5.01), (5.02, 5.01), (5.02, 5.02)]])            # covers no centre
    bbox = (0, 0, w, h)
    name, q = utils.parse_percentile_statistic(statistic)
    expected, no_cells = oracle_zonal(frame, nodata, polys, bbox, name, q)
    got, got_no_cells = geometry.aggregate.aggregate_polygons(
        to_geometries(polys), frame[np.newaxis], nodata, bbox, workloads.PROJECTION, None, name, q)
    assert got.dtype == np.float32 and got.shape == (1, len(polys))
    assert sorted(got_no_cells) == no_cells
    if name in ("sum", "mean"):
        # stated tolerance (float64 partial sums in a different order than bincount)
        np.testing.assert_allclose(got[0], expected, rtol=1e-6, equal_nan=True)
    else:
        np.testing.assert_array_equal(got[0], expected)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("dtype", ["f4", "i2"])
def test_stripe_partials_and_finalisation(dtype, resident):
    """The stripe entry points of the multi-GPU path in ONE process: the raster is cut in three row
    stripes, every stripe is reduced by gm_zonal_partials_device (only the polygons with rows in
    the stripe are visited; a resident soup keeps its preparation between calls and leaves the
    overflow flag to the finalisation), the partial vectors are added / minimised as the
    all-reduces would, gm_zonal_finalize_device forms the statistic: equal to the oracle on the
    whole raster, call after call."""
    import ctypes

    from dask_geomodeling_b200 import _native
    from dask_geomodeling_b200.geometry.aggregate import _STAT_CODES, _frame_descriptor

    h, w = 150, 130
    rng = np.random.default_rng(21)
    nodata = R.dtype_max(dtype)
    frame = (rng.uniform(0, 100, (h, w)) if dtype == "f4" else rng.integers(0, 100, (h, w))).astype(dtype)
    frame[rng.random((h, w)) < 0.1] = nodata
    polys = random_polygons(60, 150, seed=8, concave=True)
    polys.append([[(300.0, 300.0), (310.0, 300.0), (310.0, 310.0)]])      # outside the raster
    n = len(polys)
    soup = utils.PolygonSoup(to_geometries(polys))
    if resident:
        soup.to_device()
    lib = _native.lib()
    holder, nodata_ptr = _native.scalar_ptr(nodata, frame.dtype)
    bounds = [(0, 47), (47, 48), (48, 150)]                               # a one-row stripe in the middle
    stripes = [_native.DeviceArray.from_host(np.ascontiguousarray(frame[np.newaxis, a:b])) for a, b in bounds]
    sums_dev = _native.DeviceArray((3 * n,), "f8")
    extremes_dev = _native.DeviceArray((2 * n,), "f8")
    out, covered = np.empty(n, "f4"), np.empty(n, "i8")
    for statistic in ("mean", "sum", "count", "min", "max", "mean"):
        expected, no_cells = oracle_zonal(frame, nodata, polys, (0, 0, w, h), statistic)
        sums, extremes = np.zeros(3 * n), np.full(2 * n, np.finfo("f8").max)
        for (a, b), stripe in zip(bounds, stripes):
            geo = (ctypes.c_double * 6)(*utils.GeoTransform.from_bbox((0, h - b, w, h - a), b - a, w))
            polys_struct = soup.as_struct()
            desc = _frame_descriptor(stripe, 0)
            _native.check(lib.gm_zonal_partials_device(
                ctypes.byref(desc), nodata_ptr, 1, ctypes.byref(polys_struct), geo, None, 0, b - a,
                sums_dev.ptr, extremes_dev.ptr, _STAT_CODES[statistic], _native.current_stream()))
            sums += np.asarray(sums_dev)
            extremes = np.minimum(extremes, np.asarray(extremes_dev))
        totals = _native.DeviceArray.from_host(sums)
        lowest = _native.DeviceArray.from_host(extremes)
        _native.check(lib.gm_zonal_finalize_device(
            totals.ptr, lowest.ptr, n, _STAT_CODES[statistic], out.ctypes.data, covered.ctypes.data,
            ctypes.byref(polys_struct), _native.current_stream()))
        assert np.nonzero(covered == 0)[0].tolist() == no_cells
        if statistic in ("sum", "mean"):
            np.testing.assert_allclose(out, expected, rtol=1e-6, equal_nan=True)
        else:
            np.testing.assert_array_equal(out, expected)


@pytest.mark.parametrize("kind", ["uniform", "few_values", "constant", "wide_range"])
@pytest.mark.parametrize("statistic", ["median", "p90", "p1.5"])
def test_zonal_order_statistics_convex_float32(kind, statistic):
    """The shared-memory select path (float32, one span per row): value distributions that
    stress the digit placement -- few distinct values, one value, 60 binades."""
    h, w = 333, 402   # width not a multiple of 4: rows start at every 16-byte phase
    rng = np.random.default_rng(21)
    nodata = float(np.finfo("f4").max)
    if kind == "uniform":
        frame = rng.uniform(0, 100, (h, w))
    elif kind == "few_values":
        frame = rng.integers(0, 4, (h, w)).astype("f8") * 2.5 - 3
    elif kind == "constant":
        frame = np.full((h, w), 7.25)
    else:
        frame = rng.normal(0, 1, (h, w)) * 10.0 ** rng.integers(-9, 9, (h, w))
    frame = frame.astype("f4")
    frame[rng.random((h, w)) < 0.07] = nodata
    frame[40:60, 100:140] = nodata            # a polygon with no data at all below
    polys = random_polygons(70, 400, seed=8)
    polys.append([[(100.3, 40.2), (139.6, 40.2), (139.6, 59.7), (100.3, 59.7)]])
    polys.append([[(-20.5, -30.5), (450.2, -10.1), (440.7, 380.3), (-15.3, 350.9)]])  # larger than the raster
    bbox = (0, 0, w, h)
    name, q = utils.parse_percentile_statistic(statistic)
    expected, no_cells = oracle_zonal(frame, nodata, polys, bbox, name, q)
    got, got_no_cells = geometry.aggregate.aggregate_polygons(
        to_geometries(polys), frame[np.newaxis], nodata, bbox, workloads.PROJECTION, None, name, q)
    assert sorted(got_no_cells) == no_cells
    np.testing.assert_array_equal(got[0], expected)


@pytest.mark.parametrize("dtype", ["f4", "i2"])
@pytest.mark.parametrize("statistic", ["sum", "count", "max", "mean", "median", "p90"])
def test_zonal_stats_raster_without_nodata(dtype, statistic):
    """No no-data value at all: every cell under a polygon is active (the float32 paths that
    mark unused slots with the sentinel must not be taken)."""
    h, w = 97, 131
    rng = np.random.default_rng(31)
    frame = (rng.uniform(-50, 100, (h, w)) if dtype == "f4" else rng.integers(-50, 100, (h, w))).astype(dtype)
    polys = random_polygons(30, 130, seed=13, concave=True)
    bbox = (0, 0, w, h)
    name, q = utils.parse_percentile_statistic(statistic)
    label_sets = []
    for i, rings in enumerate(polys):
        labels = polyfill.burn_index([rings], bbox, h, w)
        label_sets.append((np.where(labels == 0, i, np.iinfo(np.int32).max).astype(np.int32), [i]))
    # a value no cell holds stands in for "no no-data value" in the oracle
    expected, no_cells = R.zonal_from_labels(frame, frame.dtype.type(-12345), label_sets, len(polys), name, q)
    got, got_no_cells = geometry.aggregate.aggregate_polygons(
        to_geometries(polys), frame[np.newaxis], None, bbox, workloads.PROJECTION, None, name, q)
    assert sorted(got_no_cells) == no_cells
    if name in ("sum", "mean"):
        np.testing.assert_allclose(got[0], expected, rtol=1e-6, equal_nan=True)
    else:
        np.testing.assert_array_equal(got[0], expected)


@pytest.mark.parametrize("statistic", ["mean", "max", "count", "median", "p90"])
def test_zonal_stats_frames_of_a_resident_stack(statistic):
    """Frames of a (t, h, w) stack that lives in HBM: with h * w odd the frames start at every
    16-byte phase, so the quad-aligned walks see misaligned bases, first and last lines."""
    from dask_geomodeling_b200 import _native

    t, h, w = 4, 97, 131          # 97 * 131 = 12707 cells per frame: phases 0, 3, 2, 1
    rng = np.random.default_rng(41)
    nodata = float(np.finfo("f4").max)
    stack = rng.uniform(0, 100, (t, h, w)).astype("f4")
    stack[rng.random(stack.shape) < 0.08] = nodata
    polys = random_polygons(35, 130, seed=17, concave=True)
    polys.append([[(-5.5, -7.5), (140.2, -3.1), (138.7, 101.3), (-9.3, 99.9)]])   # covers the whole raster
    bbox = (0, 0, w, h)
    name, q = utils.parse_percentile_statistic(statistic)
    resident = _native.DeviceArray.from_host(stack)
    got, _ = geometry.aggregate.aggregate_polygons(
        to_geometries(polys), resident, nodata, bbox, workloads.PROJECTION, None, name, q)
    assert got.shape == (t, len(polys))
    for frame in range(t):
        expected, _ = oracle_zonal(stack[frame], nodata, polys, bbox, name, q)
        if name == "mean":
            np.testing.assert_allclose(got[frame], expected, rtol=1e-6, equal_nan=True)
        else:
            np.testing.assert_array_equal(got[frame], expected)


def test_zonal_thresholds_and_large_polygon():
    h, w = 400, 420
    rng = np.random.default_rng(12)
    frame = rng.uniform(0, 100, (h, w)).astype("f4")
    nodata = float(np.finfo("f4").max)
    frame[rng.random((h, w)) < 0.05] = nodata
    polys = random_polygons(10, 400, seed=5)
    polys.append([[(3.3, 3.3), (410.2, 3.3), (410.2, 390.7), (3.3, 390.7)]])  # > shared memory
    thresholds = np.array([10, 50, np.nan, 0, 99, 30, 60, 5, 20, 80, 45], dtype="f4")
    for name, q in (("p90", 90.0), ("median", None), ("max", None), ("count", None)):
        stat = "percentile" if q else name
        expected, _ = oracle_zonal(frame, nodata, polys, (0, 0, w, h), stat, q, thresholds)
        got, _ = geometry.aggregate.aggregate_polygons(
            to_geometries(polys), frame[np.newaxis], nodata, (0, 0, w, h), workloads.PROJECTION,
            thresholds, stat, q)
        np.testing.assert_array_equal(got[0], expected)


def test_aggregate_raster_view_small_geometry():
    # reference tests/test_aggregate_raster.py:557-587 (centre rule + centroid fallback)
    src = raster.MemorySource(np.arange(6).reshape(2, 3).astype(float), 255, "EPSG:3857",
                              pixel_size=2.0, pixel_origin=(0, 4))
    cases = [
        ([[(2, 2), (1.9, 2), (2, 1.9)]], [3.0]),
        ([[(2, 2), (2.1, 2), (2, 1.9)]], [4.0]),
        ([[(2, 2), (2.1, 2), (2, 2.1)]], [1.0]),
        ([[(2, 2), (1.9, 2), (2, 2.1)]], [0.0]),
        ([[(2, 2), (1.9, 2), (2, 1.9)], [(2, 2), (2.1, 2), (2, 2.1)]], [3.0, 1.0]),
    ]
    request = dict(mode="intersects", projection="EPSG:3857", geometry=utils.box(0, 0, 10, 10))
    for polygons, expected in cases:
        source = geometry.MemoryGeometrySource(polygons, [{"id": i + 1} for i in range(len(polygons))])
        view = geometry.AggregateRaster(source=source, raster=src, statistic="max")
        result = view.get_data(**request)
        np.testing.assert_almost_equal(result["features"]["agg"].values, expected)
        assert result["features"]["agg"].dtype == np.float32


@pytest.mark.parametrize("dx", [0.0, 0.1, 0.4999, 0.50001, 0.9, 0.99999])
def test_aggregate_raster_no_interaction(dx):
    # reference tests/test_aggregate_raster.py:537-554
    values = np.indices((10, 10))[1][np.newaxis].astype("i4")
    src = raster.MemorySource(values, np.iinfo("i4").max, "EPSG:3857", pixel_size=1.0, pixel_origin=(0, 10))
    source = geometry.MemoryGeometrySource(
        [[(2.0 + dx, 2.0), (4.0 + dx, 2.0), (4.0 + dx, 4.0), (2.0 + dx, 4.0)],
         [(3.0, 6.0), (5, 6.0), (5, 8.0), (3, 8.0)]], [{"id": 1}, {"id": 2}])
    view = geometry.AggregateRaster(source=source, raster=src, statistic="min")
    result = view.get_data(mode="intersects", projection="EPSG:3857", geometry=utils.box(0, 0, 10, 10))
    assert result["features"]["agg"][2] == 3
