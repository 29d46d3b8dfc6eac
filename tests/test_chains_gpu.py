"""Fused chains through ``get_data`` (graph -> fusion -> one CUDA launch) against
the oracle applied block by block; configs 1 and 2 of BASELINE.json at test
sizes, plus the fusion bookkeeping (root key kept, one launch)."""
import numpy as np
import pytest

from dask_geomodeling_b200 import _native, raster, workloads
from dask_geomodeling_b200.core import fusion
from dask_geomodeling_b200._compat import config
from oracle import raster as R
from oracle import workloads as oracle_workloads

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [64, 257, 1024])
def test_cfg1_chain(size):
    a, b = workloads.cfg1_arrays(size)
    view = workloads.cfg1_view(a, b)
    before = _native.launch_count()
    got = view.get_data(**workloads.request(size, size))
    launches = _native.launch_count() - before
    values, nodata = oracle_workloads.cfg1(a, b)
    assert got["values"].dtype == values.dtype == np.uint8
    np.testing.assert_array_equal(got["values"], values)
    assert got["no_data_value"] == nodata
    # two resample kernels (float sources) + ONE fused evaluator launch
    assert launches == 3


@pytest.mark.parametrize("size", [96, 1000])
def test_cfg2_chain(size):
    ints, floats = workloads.cfg2_arrays(size, chunk=256)
    isdata, step = workloads.cfg2_views(ints, floats)
    (e_isdata, _), (e_step, e_nodata) = oracle_workloads.cfg2(ints, floats, workloads.CFG2_PAIRS)
    got = isdata.get_data(**workloads.request(size, size))
    assert got["values"].dtype == np.bool_
    np.testing.assert_array_equal(got["values"], e_isdata)
    assert got["no_data_value"] is None
    got = step.get_data(**workloads.request(size, size))
    assert got["values"].dtype == np.float32
    np.testing.assert_array_equal(got["values"], e_step)
    assert got["no_data_value"] == e_nodata


def test_fusion_keeps_root_key_and_matches_unfused():
    a, b = workloads.cfg1_arrays(128)
    view = workloads.cfg1_view(a, b)
    req = workloads.request(128, 128)
    graph, name = view.get_compute_graph(**req)
    fused = fusion.optimize(graph, name)
    assert name in fused
    assert fused[name][0] is fusion.fused_process
    assert len(fused) < len(graph)
    with config.set({"geomodeling.fuse": False}):
        unfused = view.get_data(**req)
    np.testing.assert_array_equal(view.get_data(**req)["values"], unfused["values"])


def test_window_requests_crop_and_pad():
    a, b = workloads.cfg1_arrays(64)
    view = workloads.cfg1_view(a, b)
    full = oracle_workloads.cfg1(a, b)[0]
    # window partly outside the source: outside cells are no data -> Mask fill (0)
    got = view.get_data(mode="vals", bbox=(-8, 40, 24, 72), width=32, height=32,
                        projection=workloads.PROJECTION)
    expected = np.zeros((1, 32, 32), dtype=np.uint8)
    expected[:, 8:, 8:] = full[:, 0:24, 0:24]
    np.testing.assert_array_equal(got["values"], expected)


@pytest.fixture
def small_stream_chunks(monkeypatch):
    # make the chunk pipeline kick in at test sizes: 3 streams x chunks of >= 64 rows
    monkeypatch.setattr(fusion, "STREAM_MIN_PIXELS", 1 << 12)
    monkeypatch.setattr(fusion, "STREAM_CHUNK_PIXELS", 1 << 15)


def test_streamed_chain_equals_oracle(small_stream_chunks):
    size = 517  # ragged last chunk
    ints, floats = workloads.cfg2_arrays(size, chunk=128)
    isdata, step = workloads.cfg2_views(ints, floats)
    request = workloads.request(size, size)
    graph, name = step.get_compute_graph(**request)
    assert fusion.optimize(graph, name)[name][0] is fusion.streamed_fused_process
    (e_isdata, _), (e_step, e_nodata) = oracle_workloads.cfg2(ints, floats, workloads.CFG2_PAIRS)
    got = step.get_data(**request)
    np.testing.assert_array_equal(got["values"], e_step)
    assert got["no_data_value"] == e_nodata
    np.testing.assert_array_equal(isdata.get_data(**request)["values"], e_isdata)
    with config.set({"geomodeling.stream": False}):
        np.testing.assert_array_equal(step.get_data(**request)["values"], e_step)


def test_streamed_chain_window_and_bands(small_stream_chunks):
    # three frames, request window partly outside the source (pads with no data)
    from dask_geomodeling_b200 import raster
    from oracle import raster as R

    rng = np.random.default_rng(9)
    data = rng.uniform(0, 100, (3, 300, 260)).astype("f4")
    nodata = workloads.F32_MAX
    data[rng.random(data.shape) < 0.05] = nodata
    src = raster.MemorySource(data, nodata, workloads.PROJECTION, pixel_size=1.0, pixel_origin=(0, 300),
                              time_first=0, time_delta=3600000)
    view = raster.Multiply(raster.Add(src, 1.5), src)
    request = dict(mode="vals", bbox=(-10, -20, 250, 280), width=260, height=300,
                   projection=workloads.PROJECTION, start=None, stop=None)
    from datetime import datetime

    request["start"], request["stop"] = datetime(1970, 1, 1), datetime(1970, 1, 1, 2)
    graph, name = view.get_compute_graph(**request)
    assert fusion.optimize(graph, name)[name][0] is fusion.streamed_fused_process
    window = np.full((3, 300, 260), nodata, dtype="f4")
    # row 0 of the request is y = 280 -> source row 20; column 0 is x = -10 -> 10 cells left of the source
    window[:, 0:280, 10:260] = data[:, 20:300, 0:250]
    s = R.elementwise("add", "float32", nodata, (window, nodata), 1.5)
    expected = R.elementwise("multiply", "float32", nodata, s, (window, nodata))
    got = view.get_data(**request)
    assert got["values"].shape == (3, 300, 260)
    np.testing.assert_array_equal(got["values"], expected[0])


@pytest.mark.parametrize("tile_size", [64, [100, 37], 1024])
def test_raster_tiler_equals_untiled(tile_size):
    # reference raster/parallelize.py:43-125 (tests/test_raster_parallelize.py): tiles are
    # independent requests, the stitched result equals the plain request
    size = 200
    a, b = workloads.cfg1_arrays(size)
    view = workloads.cfg1_view(a, b)
    request = workloads.request(size, size)
    request.update(bbox=(10, 20, 190, 170), width=180, height=150)
    # the oracle on the requested window: rows 200 - 170 .. 200 - 20, columns 10 .. 190
    expected, expected_nodata = oracle_workloads.cfg1(a[:, 30:180, 10:190], b[:, 30:180, 10:190])
    tiled = raster.RasterTiler(view, tile_size)
    got = tiled.get_data(**request)
    assert got["values"].dtype == expected.dtype
    np.testing.assert_array_equal(got["values"], expected)
    assert got["no_data_value"] == expected_nodata
    assert tiled.get_data(**dict(request, mode="meta")) == view.get_data(**dict(request, mode="meta"))
    # a stencil whose request margin equals its reach tiles exactly as well (Smooth does not:
    # its Gaussian reaches beyond the margin it requests, in the reference too)
    stencil = raster.MovingMax(workloads.source(a, workloads.F32_MAX), 7)
    expected, _ = R.moving_max(a[:, 27:183, 7:193], workloads.F32_MAX, 7)
    np.testing.assert_array_equal(raster.RasterTiler(stencil, [64, 50]).get_data(**request)["values"], expected)


def test_device_cache_serves_later_requests_from_hbm():
    # geomodeling.device-cache-bytes: the MemorySource arrays are uploaded once, later requests
    # (other windows included) gather from the resident copy and give the same results
    size = 300
    ints, floats = workloads.cfg2_arrays(size, chunk=128)
    isdata, step = workloads.cfg2_views(ints, floats)
    whole = workloads.request(size, size)
    window = workloads.request(size, size)
    window.update(bbox=(-20, 35, 210, 330), width=230, height=295)   # partly outside the source
    expected = [step.get_data(**whole), step.get_data(**window), isdata.get_data(**window)]
    before = _native.STATS.get("resident_hits", 0)
    with config.set({"geomodeling.device-cache-bytes": 1 << 30}):
        for _ in range(2):
            got = [step.get_data(**whole), step.get_data(**window), isdata.get_data(**window)]
            for g, e in zip(got, expected):
                assert g["values"].dtype == e["values"].dtype
                np.testing.assert_array_equal(g["values"], e["values"])
                assert g["no_data_value"] == e["no_data_value"]
    assert _native.STATS.get("resident_hits", 0) >= before + 8
