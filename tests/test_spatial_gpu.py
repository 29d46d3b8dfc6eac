"""Stencil kernels (Dilate, MovingMax, Smooth, HillShade) against the oracle, which
calls the same scipy.ndimage routines the reference calls."""
import numpy as np
from datetime import datetime as Datetime
import pytest

from dask_geomodeling_b200 import raster, workloads
from oracle import raster as R

pytestmark = pytest.mark.gpu


def dem(shape, seed=0, dtype="f4", nodata_fraction=0.02):
    rng = np.random.default_rng(seed)
    t, h, w = shape
    y, x = np.mgrid[0:h, 0:w]
    base = 50 * np.sin(x / 17.0) + 30 * np.cos(y / 11.0) + 0.05 * x + rng.normal(0, 1, (t, h, w))
    values = (base + 100).astype(dtype)
    nodata = R.dtype_max(dtype)
    values[rng.random(shape) < nodata_fraction] = nodata
    return values, nodata


@pytest.mark.parametrize("dtype", ["u1", "i2", "i4", "f4"])
@pytest.mark.parametrize("shape", [(1, 21, 34), (3, 40, 67)])
def test_dilate(dtype, shape):
    rng = np.random.default_rng(1)
    values = rng.integers(0, 6, shape).astype(dtype)
    nodata = R.dtype_max(dtype)
    wanted = [3, 1, 5]
    expected, _ = R.dilate(values, nodata, wanted)
    got = raster.Dilate.process({"values": values, "no_data_value": nodata}, wanted)
    assert got["values"].dtype == expected.dtype
    np.testing.assert_array_equal(got["values"], expected)
    assert got["no_data_value"] == nodata


@pytest.mark.parametrize("dtype", ["u1", "i2", "i4", "f4", "f8"])
@pytest.mark.parametrize("size", [3, 5, 11])
def test_moving_max(dtype, size):
    values, nodata = dem((2, 60, 83), 2, dtype=dtype, nodata_fraction=0.3)
    values[0, 20:40, 30:60] = nodata  # a hole larger than the footprint stays no data
    expected, _ = R.moving_max(values, nodata, size)
    got = raster.MovingMax.process({"values": values, "no_data_value": nodata}, size)
    assert got["values"].dtype == expected.dtype
    np.testing.assert_array_equal(got["values"], expected)


@pytest.mark.parametrize("dtype", ["f4", "f8", "i2"])
@pytest.mark.parametrize("size_px", [(5.0, 5.0), (2.0, 3.4), (6.0, 1.0)])
def test_smooth_exact(dtype, size_px):
    values, nodata = dem((2, 70, 90), 3, dtype=dtype)
    kwargs = dict(smooth_mode="exact", fill=0, size=list(size_px))
    expected, _ = R.smooth(values, nodata, size_px, 0, "exact")
    got = raster.Smooth.process({"values": values, "no_data_value": nodata}, kwargs)
    assert got["values"].dtype == expected.dtype and got["values"].shape == expected.shape
    # fp64 accumulation in scipy's tap order: expected to be bit-exact
    np.testing.assert_array_equal(got["values"], expected)


@pytest.mark.parametrize("mode,ulps", [("fma", 1), ("float32", 16)])
@pytest.mark.parametrize("dtype", ["f4", "f8"])
def test_smooth_arithmetic_modes(mode, ulps, dtype):
    """The two relaxed arithmetics of the tap sums (gm_set_smooth_mode; the default is 'fma'):
    stated tolerance 1e-6 of the value range; 'fma' differs from SciPy only
    by the rounding of the fused multiply-adds (at most the last bit of a float32 result)."""
    from dask_geomodeling_b200 import _native

    values, nodata = dem((2, 150, 333), 5, dtype=dtype)
    kwargs = dict(smooth_mode="exact", fill=0, size=[5.0, 5.0])
    expected, _ = R.smooth(values, nodata, (5.0, 5.0), 0, "exact")
    with _native.smooth_arithmetic(mode):
        got = np.asarray(raster.Smooth.process({"values": values, "no_data_value": nodata}, kwargs)["values"])
    data = values[values != nodata]
    value_range = float(data.max() - data.min())
    delta = np.abs(got.astype("f8") - expected.astype("f8"))
    assert delta.max() <= 1e-6 * value_range
    if dtype == "f4":
        assert delta.max() <= ulps * np.spacing(np.float32(np.abs(expected).max()))
    if mode == "fma" and dtype == "f4":
        assert (delta > 0).mean() < 0.05


@pytest.mark.parametrize("width", [101, 102, 130, 253])
def test_hillshade_on_a_padded_row_pitch(width):
    """Rows padded to whole 16-byte groups (what HillShade.get_sources_and_requests asks its store
    for): the quad-load path gives the very bytes of the unpadded call, and the block's own
    request equals the oracle on the unpadded window."""
    values, nodata = dem((2, 70, width), 5, dtype="f4")
    kwargs = dict(resolution=(0.5, 0.5), altitude=45.0, azimuth=315.0, fill=0)
    plain = raster.HillShade.process({"values": values, "no_data_value": nodata}, kwargs)
    pad = (-width) % 4
    padded = np.pad(values, ((0, 0), (0, 0), (0, pad)), constant_values=nodata)
    got = raster.HillShade.process({"values": padded, "no_data_value": nodata}, dict(kwargs, pad=pad))
    assert got["values"].shape == (2, 68, width - 2)
    np.testing.assert_array_equal(np.asarray(got["values"]), np.asarray(plain["values"]))
    src = raster.MemorySource(values, nodata, "EPSG:28992", 0.5, (0.0, 35.0), time_first=0, time_delta=1000)
    block = raster.HillShade(src)
    request = dict(mode="vals", bbox=(0.5, 0.5, 0.5 * (width - 1), 34.5), width=width - 2, height=68,
                   projection="EPSG:28992", start=Datetime(1970, 1, 1), stop=Datetime(1970, 1, 1, 0, 0, 1))
    through_block = block.get_data(**request)
    np.testing.assert_array_equal(through_block["values"], np.asarray(plain["values"]))


@pytest.mark.parametrize("fill", [0, 7.5])
def test_smooth_zoom(fill):
    values, nodata = dem((1, 64, 80), 4)
    size_px = [3.7, 4.2]
    kwargs = dict(smooth_mode="zoom", fill=fill, size=size_px)
    expected, _ = R.smooth(values, nodata, size_px, fill, "zoom")
    got = raster.Smooth.process({"values": values, "no_data_value": nodata}, kwargs)
    np.testing.assert_array_equal(got["values"], expected)


@pytest.mark.parametrize("dtype", ["f4", "f8", "i2"])
@pytest.mark.parametrize("angles", [(45.0, 315.0), (30.0, 100.0)])
def test_hillshade(dtype, angles):
    values, nodata = dem((2, 80, 101), 5, dtype=dtype)
    kwargs = dict(resolution=(0.5, 0.5), altitude=angles[0], azimuth=angles[1], fill=0)
    expected, expected_nodata = R.hillshade(values, nodata, kwargs["resolution"], *angles, 0)
    got = raster.HillShade.process({"values": values, "no_data_value": nodata}, kwargs)
    assert got["values"].dtype == np.uint8 and got["values"].shape == expected.shape
    assert got["no_data_value"] == expected_nodata == 256
    # stated tolerance: |delta| <= 1 grey level on at most 0.1 % of the cells
    # (float32 atan2/sin/sqrt differ in the last ulp between libm and CUDA)
    delta = np.abs(got["values"].astype(int) - expected.astype(int))
    assert delta.max() <= 1
    assert (delta > 0).mean() <= 1e-3


@pytest.mark.parametrize("dtype", ["f4", "i4"])
@pytest.mark.parametrize("size", [3, 5, 11, 15])
@pytest.mark.parametrize("pad", [0, 1, 2, 3])
def test_moving_max_padded_pitch(dtype, size, pad):
    """Windows whose rows are whole 16-byte groups (with `pad` extra columns on the right) take
    the TMA tile staging for interior tiles; the padding never enters a footprint."""
    r = size // 2
    h, w = 215, 401 + ((-(401 + 2 * r + pad)) % 4)       # (w + 2 r + pad) % 4 == 0: TMA eligible
    values, nodata = dem((2, h + 2 * r, w + 2 * r), 9, dtype=dtype, nodata_fraction=0.05)
    expected, _ = R.moving_max(values, nodata, size)
    padded = np.full((2, h + 2 * r, w + 2 * r + pad), nodata, dtype=values.dtype)
    padded[:, :, :w + 2 * r] = values
    if pad:
        padded[:, :, w + 2 * r:] = R.dtype_max(dtype) - 1       # a value that would win if it were read
    got = raster.MovingMax.process({"values": padded, "no_data_value": nodata}, size, pad)
    np.testing.assert_array_equal(np.asarray(got["values"]), expected)


def test_moving_max_view_pads_its_request():
    a, _ = workloads.cfg1_arrays(160)
    src = workloads.source(a, workloads.F32_MAX)
    req = workloads.request(131, 131)
    req["bbox"] = (10, 12, 141, 143)
    view = raster.MovingMax(src, 11)
    (_, enlarged), _, (pad, _) = view.get_sources_and_requests(**req)
    assert (enlarged["width"] * 4) % 16 == 0 and pad == enlarged["width"] - 131 - 10
    got = view.get_data(**req)
    expected = R.moving_max(a[:, 160 - 143 - 5:160 - 12 + 5, 10 - 5:141 + 5], workloads.F32_MAX, 11)[0]
    np.testing.assert_array_equal(got["values"], expected)


def test_blocks_through_get_data():
    a, _ = workloads.cfg1_arrays(96)
    src = workloads.source(a, workloads.F32_MAX)
    req = workloads.request(64, 64)
    req["bbox"] = (16, 16, 80, 80)
    for view, oracle in [
        (raster.MovingMax(src, 5), lambda v: R.moving_max(v[:, 14:82, 14:82], workloads.F32_MAX, 5)[0]),
        (raster.Smooth(src, 2.0), lambda v: R.smooth(v[:, 14:82, 14:82], workloads.F32_MAX, (2.0, 2.0), 0, "exact")[0]),
    ]:
        got = view.get_data(**req)
        np.testing.assert_array_equal(got["values"], oracle(a))


# output shapes around the kernels' tile and strip edges: one cell, widths that are not
# multiples of four, one column more / less than a strip (124) or a tile (64, 114, 128, 256)
EDGE_SHAPES = [(1, 1), (2, 3), (5, 7), (17, 123), (3, 124), (66, 125), (33, 129), (16, 257), (65, 64)]


@pytest.mark.parametrize("shape", EDGE_SHAPES)
def test_stencils_at_tile_edges(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    # HillShade: halo 1
    values, nodata = dem((1, h + 2, w + 2), 11, nodata_fraction=0.05)
    kwargs = dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0)
    expected, _ = R.hillshade(values, nodata, (1.0, 1.0), 45.0, 315.0, 0)
    got = raster.HillShade.process({"values": values, "no_data_value": nodata}, kwargs)["values"]
    assert got.shape == (1, h, w)
    delta = np.abs(got.astype(int) - expected.astype(int))
    assert delta.max() <= 1 and (delta > 0).sum() <= max(1, int(1e-3 * delta.size))
    # MovingMax 11 and 3 (float32: the four-columns-per-thread kernel), int16 (one column)
    for size, dtype in ((11, "f4"), (3, "f4"), (5, "i4"), (7, "i2")):
        r = size // 2
        values, nodata = dem((2, h + 2 * r, w + 2 * r), 12, dtype=dtype, nodata_fraction=0.3)
        expected, _ = R.moving_max(values, nodata, size)
        got = raster.MovingMax.process({"values": values, "no_data_value": nodata}, size)["values"]
        np.testing.assert_array_equal(got, expected)
    # Smooth size 5 (margin 5)
    values, nodata = dem((1, h + 10, w + 10), 13)
    expected, _ = R.smooth(values, nodata, (5.0, 5.0), 0, "exact")
    got = raster.Smooth.process({"values": values, "no_data_value": nodata},
                                dict(smooth_mode="exact", fill=0, size=[5.0, 5.0]))["values"]
    np.testing.assert_array_equal(got, expected)
    # Dilate on bytes (packed kernel) and int16 (tiled kernel), two bands
    for dtype in ("u1", "i2"):
        cells = rng.integers(0, 6, (2, h + 2, w + 2)).astype(dtype)
        expected, _ = R.dilate(cells, R.dtype_max(dtype), [3, 1, 5])
        got = raster.Dilate.process({"values": cells, "no_data_value": R.dtype_max(dtype)}, [3, 1, 5])["values"]
        np.testing.assert_array_equal(got, expected)
