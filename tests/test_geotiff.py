"""File formats either side of the path, CPU side: the GeoTIFF / VRT writer and
reader against an independent decoder (Pillow's libtiff), and the reference's sink tests
(`dask_geomodeling/tests/test_raster_sinks.py`) replayed on RasterFileSink / to_file with a CPU
mock source (the sink itself is host I/O; tests/test_files_gpu.py covers the device side)."""
import os
import struct
import zlib
from datetime import datetime, timedelta
from unittest.mock import patch

import numpy as np
import pytest

from dask_geomodeling_b200 import geotiff
from dask_geomodeling_b200._compat import config
from dask_geomodeling_b200.core import Block
from dask_geomodeling_b200.raster import RasterFileSink, RasterTiler, to_file

from mocks import MockRaster

PIL_Image = pytest.importorskip("PIL.Image")

GT = (136700.0, 5.0, 0.0, 455800.0, 0.0, -5.0)


@pytest.fixture
def root(tmp_path):
    with config.set({"geomodeling.root": str(tmp_path)}):
        yield str(tmp_path)


@pytest.mark.parametrize("dtype", ["u1", "i2", "u2", "i4", "f4", "f8"])
@pytest.mark.parametrize("shape", [(1, 1), (37, 300), (256, 256), (513, 260)])
def test_round_trip_and_independent_decoder(tmp_path, dtype, shape):
    rng = np.random.default_rng(3)
    values = (rng.uniform(0, 200, shape)).astype(dtype)
    path = str(tmp_path / "a.tif")
    geotiff.write_geotiff(path, values, GT, "EPSG:28992", no_data_value=255)
    tif = geotiff.GeoTiff(path)
    assert tif.shape == (1,) + shape and tif.dtype == np.dtype(dtype)
    assert tif.geo_transform == GT and tif.projection == "EPSG:28992" and tif.no_data_value == 255.0
    assert (tif.block_w, tif.block_h, tif.compression) == (256, 256, 8)
    np.testing.assert_array_equal(tif.read()[0], values)
    np.testing.assert_array_equal(tif.read_window(0, 1, shape[0] // 3, shape[0], shape[1] // 2, shape[1])[0],
                                  values[shape[0] // 3:, shape[1] // 2:])
    if dtype != "f8":       # (Pillow has no float64 image mode)
        with PIL_Image.open(path) as image:
            np.testing.assert_array_equal(np.asarray(image), values)
            tags = image.tag_v2
            assert tags[33550] == (5.0, 5.0, 0.0) and tags[33922] == (0.0, 0.0, 0.0, 136700.0, 455800.0, 0.0)
            assert tags[42113] == "255"
            assert tuple(tags[34735]) == (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 28992)


def test_bands_wkt_and_geographic(tmp_path):
    values = np.arange(3 * 40 * 50, dtype="i2").reshape(3, 40, 50)
    path = str(tmp_path / "b.tif")
    wkt = 'PROJCS["custom",GEOGCS["x",DATUM["d",SPHEROID["s",6378137,298.25]]]]'
    geotiff.write_geotiff(path, values, GT, wkt, no_data_value=-1, compress=False)
    tif = geotiff.GeoTiff(path)
    assert tif.bands == 3 and tif.projection == wkt and tif.no_data_value == -1.0 and tif.compression == 1
    np.testing.assert_array_equal(tif.read(), values)
    np.testing.assert_array_equal(tif.read_window(1, 3, 5, 9, 7, 50), values[1:3, 5:9, 7:50])
    geotiff.write_geotiff(path, values[:1], (4.0, 0.5, 0, 52.0, 0, -0.5), "epsg:4326")
    tif = geotiff.GeoTiff(path)
    assert tif.projection == "EPSG:4326" and tif.no_data_value is None
    assert tif._geo_key(1024) == 2 and tif._geo_key(2048) == 4326


@pytest.mark.parametrize("dtype,predictor", [("i2", 2), ("u1", 2), ("f4", 2), ("f4", 3), ("f8", 3)])
def test_predictors(tmp_path, dtype, predictor):
    """Horizontal differencing and the floating-point predictor (what GDAL writes with PREDICTOR=2 / 3):
    written here, read back here and -- where Pillow has an image mode -- by libtiff."""
    y, x = np.mgrid[0:300, 0:270]
    values = (np.sin(x / 30.0) * 100 + y * 0.25).astype(dtype)
    path = str(tmp_path / "p.tif")
    geotiff.write_geotiff(path, values, GT, "EPSG:28992", None, predictor=predictor)
    plain = str(tmp_path / "q.tif")
    geotiff.write_geotiff(plain, values, GT, "EPSG:28992", None)
    assert os.path.getsize(path) < os.path.getsize(plain)         # the predictor pays on smooth data
    tif = geotiff.GeoTiff(path)
    assert tif.predictor == predictor
    np.testing.assert_array_equal(tif.read()[0], values)
    np.testing.assert_array_equal(tif.read_window(0, 1, 250, 300, 200, 270)[0], values[250:, 200:])
    if dtype != "f8":
        with PIL_Image.open(path) as image:
            np.testing.assert_array_equal(np.asarray(image), values)
    with pytest.raises(ValueError):
        geotiff.write_geotiff(path, values.astype("i4"), GT, "EPSG:28992", None, predictor=3)


def test_bigtiff_layout(tmp_path):
    """The 64-bit layout (files past 4 GB) forced on a small raster: offsets and counts are 8-byte
    words, the directory has 20-byte entries; our reader and libtiff read it back."""
    values = np.arange(300 * 520, dtype="f4").reshape(300, 520)
    path = str(tmp_path / "big.tif")
    geotiff.write_geotiff(path, values, GT, "EPSG:28992", -9999.0, bigtiff=True)
    with open(path, "rb") as f:
        assert struct.unpack("<2sHHH", f.read(8)) == (b"II", 43, 8, 0)
    tif = geotiff.GeoTiff(path)
    assert tif._big and tif.geo_transform == GT and tif.no_data_value == -9999.0
    np.testing.assert_array_equal(tif.read()[0], values)
    np.testing.assert_array_equal(tif.read_window(0, 1, 250, 300, 500, 520)[0], values[250:, 500:])
    with PIL_Image.open(path) as image:
        np.testing.assert_array_equal(np.asarray(image), values)


def test_reads_files_of_other_writers(tmp_path):
    """Strips, big-endian order, the horizontal predictor and chunky samples do not come out of
    our writer: a strip file written by Pillow and hand-made files cover the reader."""
    values = (np.arange(70 * 90) % 251).astype("u1").reshape(70, 90)
    path = str(tmp_path / "pil.tif")
    PIL_Image.fromarray(values).save(path, compression="tiff_adobe_deflate")
    tif = geotiff.GeoTiff(path)
    assert 322 not in tif.tags and tif.compression == 8
    np.testing.assert_array_equal(tif.read()[0], values)
    np.testing.assert_array_equal(tif.read_window(0, 1, 60, 70, 3, 11)[0], values[60:70, 3:11])

    # big-endian, two chunky int16 samples, predictor 2, two strips of 2 rows
    data = np.array([[[1, -5], [4, 7], [300, -300]], [[9, 9], [8, 8], [7, 7]],
                     [[0, 1], [2, 3], [4, 5]], [[-1, -2], [-3, -4], [-5, -6]]], dtype=">i2")   # (4, 3, 2)
    diff = data.copy()
    diff[:, 1:] = data[:, 1:] - data[:, :-1]
    strips = [zlib.compress(diff[:2].tobytes()), zlib.compress(diff[2:].tobytes())]
    body = b"".join(strips)
    tags = [(256, 3, 1, 3), (257, 3, 1, 4), (258, 3, 2, None), (259, 3, 1, 8), (262, 3, 1, 1),
            (273, 4, 2, None), (277, 3, 1, 2), (278, 3, 1, 2), (279, 4, 2, None), (284, 3, 1, 1),
            (317, 3, 1, 2), (339, 3, 2, None)]
    ifd_at = 8 + len(body)
    extra_at = ifd_at + 2 + 12 * len(tags) + 4
    extra = struct.pack(">2I", 8, 8 + len(strips[0])) + struct.pack(">2I", len(strips[0]), len(strips[1]))
    ifd = struct.pack(">H", len(tags))
    for tag, kind, count, value in tags:
        ifd += struct.pack(">HHI", tag, kind, count)
        if tag == 258:
            ifd += struct.pack(">HH", 16, 16)
        elif tag == 339:
            ifd += struct.pack(">HH", 2, 2)
        elif tag == 273:
            ifd += struct.pack(">I", extra_at)
        elif tag == 279:
            ifd += struct.pack(">I", extra_at + 8)
        else:
            ifd += struct.pack(">HH", value, 0)
    path = str(tmp_path / "mm.tif")
    with open(path, "wb") as f:
        f.write(struct.pack(">2sHI", b"MM", 42, ifd_at) + body + ifd + struct.pack(">I", 0) + extra)
    tif = geotiff.GeoTiff(path)
    assert tif.shape == (2, 4, 3) and tif.dtype == np.dtype("i2") and tif.planar == 1
    np.testing.assert_array_equal(tif.read(), np.moveaxis(data.astype("i2"), 2, 0))
    np.testing.assert_array_equal(tif.read_window(1, 2, 1, 4, 1, 3)[0], data[1:4, 1:3, 1])
    assert tif.geo_transform is None and tif.projection is None


@pytest.mark.parametrize("dtype,mode", [("u1", None), ("i4", "I"), ("f4", "F")])
def test_reads_lzw_files(tmp_path, dtype, mode):
    """LZW strips written by libtiff (Pillow): smooth data (long matches, the table fills and is
    cleared) and noise (mostly literals, code widths up to 12 bits)."""
    rng = np.random.default_rng(4)
    y, x = np.mgrid[0:333, 0:411]
    for name, values in (("smooth", ((x // 7 + y // 5) % 200).astype(dtype)),
                         ("noise", rng.integers(0, 250, (333, 411)).astype(dtype))):
        path = str(tmp_path / (name + ".tif"))
        PIL_Image.fromarray(values, mode=mode).save(path, compression="tiff_lzw")
        tif = geotiff.GeoTiff(path)
        assert tif.compression == 5 and tif.dtype == np.dtype(dtype)
        np.testing.assert_array_equal(tif.read()[0], values)
        np.testing.assert_array_equal(tif.read_window(0, 1, 300, 333, 7, 400)[0], values[300:, 7:400])


def test_unsupported_files_raise(tmp_path):
    path = str(tmp_path / "c.tif")
    PIL_Image.fromarray(np.zeros((4, 4), "u1")).save(path, compression="packbits")
    with pytest.raises(NotImplementedError):
        geotiff.GeoTiff(path)
    with open(path, "wb") as f:
        f.write(b"not a tiff at all")
    with pytest.raises(IOError):
        geotiff.open_raster(path)


def test_vrt_mosaic(tmp_path):
    full = np.arange(8 * 12, dtype="f4").reshape(8, 12)
    for k, (r, c) in enumerate([(0, 0), (0, 6), (4, 6)]):      # the tile at (4, 0) is missing
        gt = (100.0 + c * 2.0, 2.0, 0, 50.0 - r * 2.0, 0, -2.0)
        geotiff.write_geotiff(str(tmp_path / "t{}.tif".format(k)), full[r:r + 4, c:c + 6], gt, "EPSG:3857", -9999.0)
    target = str(tmp_path / "m.vrt")
    geotiff.write_vrt(target, [str(tmp_path / "t{}.tif".format(k)) for k in range(3)])
    mosaic = geotiff.open_raster(target)
    assert mosaic.shape == (1, 8, 12) and mosaic.dtype == np.dtype("f4")
    assert mosaic.geo_transform == (100.0, 2.0, 0.0, 50.0, 0.0, -2.0)
    assert mosaic.projection == "EPSG:3857" and mosaic.no_data_value == -9999.0
    expected = full.copy()
    expected[4:, :6] = -9999.0
    np.testing.assert_array_equal(mosaic.read()[0], expected)
    np.testing.assert_array_equal(mosaic.read_window(0, 1, 2, 7, 3, 9)[0], expected[2:7, 3:9])


# ---- reference tests/test_raster_sinks.py -------------------------------------------------------

@pytest.fixture
def source():
    return MockRaster(origin=datetime(2000, 1, 1), timedelta=timedelta(hours=1), bands=1,
                      value=np.full((100, 100), 7, dtype=np.uint8), projection="EPSG:3857")


@pytest.fixture
def request_kwargs():
    return {"mode": "vals", "bbox": (0, 0, 100, 100), "projection": "EPSG:3857", "width": 4, "height": 4,
            "start": datetime(2000, 1, 1), "stop": datetime(2000, 1, 1)}


@pytest.fixture
def tiled_output(source, root, request_kwargs):
    path = os.path.join(root, "tiled_output")
    RasterTiler(RasterFileSink(source, path), 2).get_data(**dict(request_kwargs, width=8, height=8))
    return path


def test_sink_init(source, root):
    path = os.path.join(root, "test_init")
    sink = RasterFileSink(source, path)
    assert sink.store is source and sink.url == "file://" + path
    assert RasterFileSink(source, "relative").url == "file://" + os.path.join(root, "relative")
    with pytest.raises(TypeError):
        RasterFileSink("not_a_raster", path)
    with pytest.raises(NotImplementedError):
        RasterFileSink(source, "http://example.com/x")


def test_strict_file_paths(source, root, tmp_path_factory):
    from dask_geomodeling_b200 import utils

    outside = str(tmp_path_factory.mktemp("elsewhere") / "x")
    assert utils.safe_abspath(outside) == outside          # absolute paths pass by default
    with config.set({"geomodeling.strict-file-paths": True}):
        assert utils.safe_file_url("inside/a.tif") == "file://" + os.path.join(root, "inside", "a.tif")
        with pytest.raises(IOError):
            RasterFileSink(source, outside)


def test_sink_process(source, root, request_kwargs):
    path = os.path.join(root, "sink")
    assert RasterFileSink(source, path).get_data(**request_kwargs) is None
    files = [f for f in os.listdir(path) if f.endswith(".tif")]
    assert len(files) == 1
    tif = geotiff.GeoTiff(os.path.join(path, files[0]))
    assert (tif.width, tif.height) == (4, 4)
    assert (tif.read()[0] == 7).all() and tif.no_data_value == 255.0
    assert tif.geo_transform == pytest.approx((0, 25, 0, 100, 0, -25))
    with PIL_Image.open(os.path.join(path, files[0])) as image:
        assert (np.asarray(image) == 7).all()


@pytest.mark.parametrize("overrides", [{"start": datetime(2099, 1, 1), "stop": datetime(2099, 1, 1)},
                                       {"bbox": (1000, 1000, 1100, 1100)}])
def test_sink_no_data_creates_no_files(source, root, request_kwargs, overrides):
    path = os.path.join(root, "nothing")
    assert RasterFileSink(source, path).get_data(**{**request_kwargs, **overrides}) is None
    assert not os.path.exists(path)


def test_sink_non_vals_mode_forwards(source, root, request_kwargs):
    result = RasterFileSink(source, os.path.join(root, "t")).get_data(**dict(request_kwargs, mode="time"))
    assert len(result["time"]) == 1


def test_sink_rejects_multi_band(root):
    data = {"values": np.zeros((2, 3, 3), "u1"), "no_data_value": 255}
    with pytest.raises(ValueError):
        RasterFileSink.process(data, {"url": "file://" + root, "hash": "x", "bbox": (0, 0, 1, 1), "projection": "EPSG:3857"})


def test_tiled_sink_and_merge(tiled_output, root):
    assert len([f for f in os.listdir(tiled_output) if f.endswith(".tif")]) == 16
    target = os.path.join(root, "merged.vrt")
    RasterFileSink.merge_files(tiled_output, target)
    mosaic = geotiff.open_raster(target)
    assert (mosaic.width, mosaic.height) == (8, 8) and (mosaic.read() == 7).all()
    with pytest.raises(IOError):
        RasterFileSink.merge_files(tiled_output, target)
    os.makedirs(os.path.join(root, "empty_dir"))
    with pytest.raises(IOError):
        RasterFileSink.merge_files(os.path.join(root, "empty_dir"), os.path.join(root, "none.vrt"))


def test_to_file(source, root, request_kwargs):
    kwargs = {k: v for k, v in request_kwargs.items() if k != "mode"}
    target = os.path.join(root, "to_file_output.vrt")
    to_file(source, target, tile_size=2, **kwargs)
    mosaic = geotiff.open_raster(target)
    assert mosaic.shape == (1, 4, 4) and (mosaic.read() == 7).all()
    target = os.path.join(root, "sub", "block_to_file.vrt")
    os.makedirs(os.path.dirname(target))
    source.to_file(target, tile_size=2, **kwargs)     # RasterBlock.to_file
    assert geotiff.open_raster(target).shape == (1, 4, 4)


def test_to_file_auto_defaults(source, root, request_kwargs):
    class WithGrid(MockRaster):
        geometry = property(lambda self: type("G", (), {"bounds": (0, 0, 100, 100)})())
        geo_transform = (0, 1, 0, 100, 0, -1)

    grid_source = WithGrid(origin=datetime(2000, 1, 1), timedelta=timedelta(hours=1), bands=1,
                           value=np.full((100, 100), 7, dtype=np.uint8), projection="EPSG:3857")
    with patch.object(Block, "get_data") as get_data, patch.object(RasterFileSink, "merge_files"):
        to_file(grid_source, os.path.join(root, "auto.vrt"), tile_size=50,
                start=request_kwargs["start"], stop=request_kwargs["stop"])
    request = get_data.call_args[1]
    assert request["projection"] == "EPSG:3857" and request["bbox"] == (0, 0, 100, 100)
    assert request["width"] == 100 and request["height"] == 100
    with pytest.raises(ValueError):     # the plain mock has no geometry
        to_file(source, os.path.join(root, "auto2.vrt"), tile_size=50)
