"""File formats either side of the path on the device side: RasterFileSource feeds the
CUDA path (window decode -> upload -> nearest-neighbour gather), RasterFileSink / to_file take
tiles that were evaluated on the device.  Reference tests replayed: TestGeoTIFFSource
(`dask_geomodeling/tests/test_raster_sources.py:272-300` over the shared cases `:68-213`) and
`tests/test_raster_sinks.py:166-178`."""
import os
from datetime import datetime, timedelta

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from dask_geomodeling_b200 import geotiff, raster
from dask_geomodeling_b200._compat import config
from dask_geomodeling_b200.raster import MemorySource, RasterFileSource

from test_reference_replay_gpu import _get, _nearest_neighbour

pytestmark = pytest.mark.gpu

GT = (136700.0, 5.0, 0.0, 455800.0, 0.0, -5.0)


@pytest.fixture
def root(tmp_path):
    with config.set({"geomodeling.root": str(tmp_path)}):
        yield str(tmp_path)


@pytest.fixture
def single_pixel(root):
    geotiff.write_geotiff(os.path.join(root, "single_pixel.tif"), np.array([[[5]]], "u1"), GT, "EPSG:28992", 255)
    return RasterFileSource(url="single_pixel.tif")


@pytest.fixture
def temporal(root):
    geotiff.write_geotiff(os.path.join(root, "test_temporal.tif"), np.array([[[4]], [[5]]], "u1"), GT, "EPSG:28992", 255)
    return RasterFileSource(url=os.path.join(root, "test_temporal.tif"), time_first=datetime(2000, 1, 1),
                            time_delta=timedelta(days=1))


def test_attributes(single_pixel, temporal, root):
    assert single_pixel.url == "file://" + os.path.join(root, "single_pixel.tif")
    assert single_pixel.dtype == np.dtype("u1") and single_pixel.fillvalue == 255
    assert single_pixel.projection == "EPSG:28992" and tuple(single_pixel.geo_transform) == GT
    assert len(single_pixel) == 1 and not single_pixel.temporal and single_pixel.timedelta is None
    assert single_pixel.period == (datetime(1970, 1, 1),) * 2
    assert single_pixel.geometry.bounds == (136700.0, 455795.0, 136705.0, 455800.0)
    assert len(temporal) == 2 and temporal.temporal and temporal.timedelta == timedelta(days=1)
    assert temporal.period == (datetime(2000, 1, 1), datetime(2000, 1, 2))
    single_pixel.close_dataset()
    assert single_pixel.dtype == np.dtype("u1")     # reopened on demand


def test_point_and_bbox_requests(single_pixel):
    for dx, dy in ((0, 0), (0, -4.99), (4.99, 0), (4.99, -4.99)):
        data = _get(single_pixel, (136700 + dx, 455800 + dy) * 2, 1, 1)
        assert data["values"].shape == (1, 1, 1) and data["values"][0, 0, 0] == 5
    for dx, dy in ((0, -5.0), (5.0, 0), (-5.0, 5.0), (-0.01, 0), (0, 0.01)):
        data = _get(single_pixel, (136700 + dx, 455800 + dy) * 2, 1, 1)
        assert data["values"][0, 0, 0] == data["no_data_value"] == 255
    assert _get(single_pixel, (136700, 455795, 136705, 455800), 1, 1)["values"].tolist() == [[[5]]]
    for dx, dy in ((0, -5), (-5, 0), (0, 5), (5, 0)):
        data = _get(single_pixel, (136700 + dx, 455795 + dy, 136705 + dx, 455800 + dy), 1, 1)
        assert data["values"].tolist() == [[[255]]]
    assert _get(single_pixel, (136700, 455795, 136710, 455800), 4, 2)["values"].tolist() == \
        [[[5, 5, 255, 255], [5, 5, 255, 255]]]
    assert _get(single_pixel, (136700, 455790, 136705, 455800), 1, 2)["values"].tolist() == [[[5], [255]]]


def test_time_and_band_selection(temporal):
    request = dict(mode="vals", projection="EPSG:28992", bbox=(136700, 455795, 136705, 455800), width=1, height=1)
    assert temporal.get_data(**request)["values"].tolist() == [[[5]]]       # no start: the last frame
    both = temporal.get_data(start=datetime(2000, 1, 1), stop=datetime(2000, 1, 2), **request)
    assert both["values"].tolist() == [[[4]], [[5]]]
    assert temporal.get_data(start=datetime(2000, 1, 1), **request)["values"].tolist() == [[[4]]]
    assert temporal.get_data(start=datetime(2001, 1, 1), stop=datetime(2001, 2, 1), **request) is None
    times = temporal.get_data(mode="time", start=datetime(2000, 1, 1), stop=datetime(2000, 1, 2))
    assert times["time"] == [datetime(2000, 1, 1), datetime(2000, 1, 2)]
    assert temporal.get_data(mode="meta", start=datetime(2000, 1, 1), stop=datetime(2000, 1, 2))["meta"] == [None, None]
    assert temporal.get_data(mode="meta", start=datetime(1970, 1, 1), stop=datetime(1971, 1, 1))["meta"] == []


@pytest.mark.parametrize("dtype", ["u1", "i2", "f4", "f8"])
def test_window_requests_equal_the_memory_source(root, dtype):
    """Aligned, shifted, zoomed and partly-outside requests on a 3-frame 600 x 530 file (several
    256-cell tiles): equal to the restated GDAL nearest-neighbour warp and to MemorySource."""
    rng = np.random.default_rng(11)
    data = rng.integers(0, 200, (3, 600, 530)).astype(dtype)
    nodata, origin, cell = 250, (1000.0, 2000.0), 2.5
    geotiff.write_geotiff(os.path.join(root, "big.tif"), data, (origin[0], cell, 0, origin[1], 0, -cell), "EPSG:28992", nodata)
    src = RasterFileSource("big.tif", time_first=0, time_delta=3600000)
    mem = MemorySource(data, nodata, "EPSG:28992", cell, origin, time_first=0, time_delta=3600000)
    requests = [
        ((1000.0, 2000.0 - 600 * cell, 1000.0 + 530 * cell, 2000.0), 600, 530),
        ((1640.0, 1200.0, 1740.0, 1400.0), 80, 40),             # aligned window over a tile corner
        ((1003.1, 1421.7, 1917.3, 1998.2), 64, 96),
        ((1630.3, 1330.2, 1700.9, 1380.6), 133, 141),           # zoom in
        ((880.3, 390.2, 2470.9, 2130.6), 33, 41),               # larger than the file
        ((5000.0, 5000.0, 5100.0, 5100.0), 4, 4),               # nothing of the file
    ]
    start, stop = datetime(1970, 1, 1), datetime(1970, 1, 1, 2)
    for bbox, height, width in requests:
        kwargs = dict(mode="vals", projection="EPSG:28992", bbox=bbox, width=width, height=height, start=start, stop=stop)
        got = src.get_data(**kwargs)
        assert got["values"].dtype == data.dtype and got["no_data_value"] == nodata
        assert_array_equal(got["values"], _nearest_neighbour(data, nodata, origin, cell, bbox, height, width))
        assert_array_equal(got["values"], mem.get_data(**kwargs)["values"])


def test_file_source_feeds_a_fused_chain_and_to_file_round_trips(root):
    """file -> (x * 2 + 1 clipped by a mask) on the device -> to_file -> file: the VRT read back
    equals the same chain on a MemorySource."""
    rng = np.random.default_rng(2)
    data = rng.uniform(0, 100, (1, 300, 420)).astype("f4")
    data[0, 10:40, 50:90] = -9999.0
    gt = (0.0, 2.0, 0, 600.0, 0, -2.0)
    geotiff.write_geotiff(os.path.join(root, "in.tif"), data, gt, "EPSG:28992", -9999.0)
    src = RasterFileSource("in.tif")
    view = raster.Clip(src * 2.0 + 1.0, src > 20.0)
    request = dict(bbox=(0.0, 0.0, 840.0, 600.0), width=420, height=300, projection="EPSG:28992")
    view.to_file(os.path.join(root, "out.vrt"), tile_size=[256, 128], **request)
    assert len(os.listdir(os.path.join(root, "tiles"))) == 2 * 3
    mem = MemorySource(data, -9999.0, "EPSG:28992", 2.0, (0.0, 600.0))
    expected = raster.Clip(mem * 2.0 + 1.0, mem > 20.0).get_data(mode="vals", **request)
    back = RasterFileSource(os.path.join(root, "out.vrt"))
    assert back.dtype == np.dtype("f4") and back.projection == "EPSG:28992"
    assert back.fillvalue == np.float32(expected["no_data_value"])
    got = back.get_data(mode="vals", **request)
    assert_array_equal(got["values"], expected["values"])
    assert got["no_data_value"] == expected["no_data_value"]
