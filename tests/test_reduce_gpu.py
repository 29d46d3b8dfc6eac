"""Max / Group / reduce_rasters through the blocks' ``get_data``; the golden
cases of the reference's own functions are replayed by tests/test_golden_gpu.py."""
from datetime import datetime, timedelta

import numpy as np
import pytest

from dask_geomodeling_b200 import raster, workloads
from dask_geomodeling_b200.raster.reduction import reduce_rasters
from oracle import raster as R

pytestmark = pytest.mark.gpu


def sources(n, size=48, frames=1, seed=3, dtype="f4", **kwargs):
    rng = np.random.default_rng(seed)
    nodata = R.dtype_max(dtype)
    out = []
    for _ in range(n):
        values = rng.uniform(0, 100, (frames, size, size)).astype(dtype)
        values[rng.random(values.shape) < 0.4] = nodata
        out.append((values, nodata))
    return out


def test_max_view_and_fusion():
    (a, nd), (b, _), (c, _) = sources(3)
    sa, sb, sc = (workloads.source(x, nd) for x in (a, b, c))
    request = workloads.request(48, 48)
    got = raster.Max(sa, sb, sc).get_data(**request)
    expected, _ = R.reduce_rasters([(a, nd), (b, nd), (c, nd)], "max", nd, "float32")
    np.testing.assert_array_equal(got["values"], expected)
    assert got["no_data_value"] == nd
    # fused with element-wise blocks on both sides: Max(Add(a, b), c) * 2
    view = raster.Multiply(raster.Max(raster.Add(sa, sb), sc), 2.0)
    got = view.get_data(**request)
    added, _ = R.elementwise("add", "float32", nd, (a, nd), (b, nd))
    best, _ = R.reduce_rasters([(added, nd), (c, nd)], "max", nd, "float32")
    expected, _ = R.elementwise("multiply", "float32", nd, (best, nd), 2.0)
    np.testing.assert_array_equal(got["values"], expected)


def test_max_validates_and_handles_missing_data():
    (a, nd), = sources(1)
    with pytest.raises(TypeError):
        raster.Max(workloads.source(a, nd), 3)
    sa = workloads.source(a, nd)
    payload = {"values": a, "no_data_value": nd}
    assert raster.Max.process({"dtype": "float32", "fillvalue": nd}, None, None) is None
    got = raster.Max.process({"dtype": "float32", "fillvalue": nd}, None, payload)
    np.testing.assert_array_equal(np.asarray(got["values"]), a)
    assert raster.Max.process({"dtype": "float32", "fillvalue": nd}, {"time": [1]}, payload) == {"time": [1]}
    with pytest.raises(NotImplementedError):
        reduce_rasters([payload, payload], "median")
    with pytest.raises(KeyError):
        reduce_rasters([payload], "nonsense")
    with pytest.raises(ValueError):
        reduce_rasters([], "max")
    assert raster.Max(sa, sa).dtype == np.float32


def test_group_of_aligned_temporal_sources():
    """Two equidistant sources whose periods overlap by one frame: merged by bands."""
    (a, nd), (b, _) = sources(2, frames=3, seed=9)
    hour = 3600 * 1000
    sa = workloads.source(a, nd, time_first=0, time_delta=hour)
    sb = workloads.source(b, nd, time_first=2 * hour, time_delta=hour)
    view = raster.Group(sa, sb)
    assert view.timedelta == timedelta(hours=1)
    assert view.period == (datetime(1970, 1, 1), datetime(1970, 1, 1, 4))
    request = workloads.request(48, 48, start=datetime(1970, 1, 1), stop=datetime(1970, 1, 1, 4))
    got = view.get_data(**request)
    expected, fill = R.group_by_bands([(a, nd), (b, nd)], [(0, 3), (2, 5)], "f4", (5, 48, 48))
    np.testing.assert_array_equal(got["values"], expected)
    assert got["no_data_value"] == fill
    assert view.get_data(mode="time", start=request["start"], stop=request["stop"])["time"] == [
        datetime(1970, 1, 1, h) for h in range(5)]
    # the latest frame only
    got = view.get_data(**workloads.request(48, 48))
    np.testing.assert_array_equal(got["values"], expected[4:5])


def test_group_of_non_temporal_sources_fills_gaps():
    (a, nd), (b, _) = sources(2, seed=11)
    view = raster.Group(workloads.source(a, nd), workloads.source(b, nd))
    got = view.get_data(**workloads.request(48, 48))
    expected, _ = R.reduce_rasters([(a, nd), (b, nd)], "last", nd, "float32")
    np.testing.assert_array_equal(got["values"], expected)


def test_place_through_get_data():
    """Place as a view (reference: raster/spatial.py:440-731): the source placed at three
    coordinates, overlapping copies merged with 'last' and 'max'; against the oracle."""
    rng = np.random.default_rng(4)
    data = rng.uniform(1, 100, (1, 6, 8)).astype("f4")
    nodata = workloads.F32_MAX
    data[rng.random(data.shape) < 0.3] = nodata
    src = raster.MemorySource(data, nodata, workloads.PROJECTION, pixel_size=1.0, pixel_origin=(10, 26))
    coordinates = [(5.0, 5.0), (12.0, 10.0), (14.0, 6.0)]
    request = dict(mode="vals", bbox=(0, 0, 30, 20), width=30, height=20, projection=workloads.PROJECTION)
    for statistic in ("last", "max"):
        view = raster.Place(src, workloads.PROJECTION, (12.0, 22.0), coordinates, statistic)
        got = view.get_data(**request)
        kwargs = {"anchor": (12.0, 22.0), "src_bbox": (10.0, 20.0, 18.0, 26.0), "dst_bbox": (0.0, 0.0, 30.0, 20.0),
                  "cellsize": (1.0, 1.0), "statistic": statistic, "coordinates": coordinates}
        expected, _ = R.place_warp(data, nodata, kwargs)
        np.testing.assert_array_equal(got["values"], expected)
    # a request finer than the source: one shifted request per coordinate ("group" mode)
    fine = dict(mode="vals", bbox=(4, 4, 8, 8), width=2, height=2, projection=workloads.PROJECTION)
    view = raster.Place(src, workloads.PROJECTION, (12.0, 22.0), [(5.0, 5.0)], "last")
    plan = view.get_sources_and_requests(**fine)
    assert plan[0][0]["mode"] == "group" and len(plan) == 2
    assert view.get_data(**fine)["values"].shape == (1, 2, 2)
    with pytest.raises(ValueError):
        raster.Place(src, workloads.PROJECTION, (1, 2, 3), coordinates)
    with pytest.raises(ValueError):
        raster.Place(src, workloads.PROJECTION, (1, 2), coordinates, "nonsense")
