"""BASELINE.json configurations at their full sizes, checked through properties that do not
need a full-size CPU run: closed forms, window / stripe invariance, sequential-in-t
accumulation on the device, ordering of order statistics, and the oracle on sampled windows.
Inputs of configs 3-5 are generated in HBM (torch) and handed to the blocks' ``process``
functions as device arrays, exactly as the benchmarks do."""
import numpy as np
import pytest

from dask_geomodeling_b200 import _native, geometry, raster, utils, workloads
from dask_geomodeling_b200.core import fusion
from oracle import polyfill
from oracle import raster as R
from oracle import workloads as oracle_workloads

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
NODATA = workloads.F32_MAX


def wrap(t):
    dtype = {torch.float32: "f4", torch.uint8: "u1", torch.int32: "i4"}[t.dtype]
    return _native.DeviceArray(tuple(t.shape), dtype, ptr=t.data_ptr(), owner=t)


def on_device(fn):
    torch.cuda.synchronize()   # torch filled the inputs on its own stream
    with fusion.device_resident():
        out = fn()
    _native.synchronize()
    return out


# ---- config 2: Reclassify + Clip + Step + IsData, 16384 x 16384 -----------------------------------
def test_cfg2_full_size_closed_form_and_stripes():
    size = 16384
    ints, floats = workloads.cfg2_arrays(size)
    isdata, step = workloads.cfg2_views(ints, floats)
    got = isdata.get_data(**workloads.request(size, size))["values"]
    # closed form of the chain: a cell has data iff the float raster has data and the class is a key
    keys = np.array([k for k, _ in workloads.CFG2_PAIRS], dtype=np.int16)
    expected = (floats != np.float32(NODATA)) & np.isin(ints, keys)
    assert got.dtype == np.bool_ and np.array_equal(got, expected)
    # the oracle on three 256-row stripes, and the same stripes requested on their own
    for r0 in (0, 7777, size - 256):
        r1 = r0 + 256
        (e_isdata, _), (e_step, _) = oracle_workloads.cfg2(ints[:, r0:r1], floats[:, r0:r1], workloads.CFG2_PAIRS)
        assert np.array_equal(got[:, r0:r1], e_isdata)
        request = workloads.request(size, size)
        request.update(bbox=(0, size - r1, size, size - r0), height=256)
        assert np.array_equal(step.get_data(**request)["values"], e_step)


# ---- config 3: Smooth / MovingMax / HillShade on a 32768 x 32768 DEM ---------------------------------
@pytest.fixture(scope="module")
def dem32k():
    n = 32768
    gen = torch.Generator(device="cuda").manual_seed(3)
    z = torch.empty((1, n, n), dtype=torch.float32, device="cuda")
    y = torch.arange(n, device="cuda", dtype=torch.float32)[:, None]
    x = torch.arange(n, device="cuda", dtype=torch.float32)[None, :]
    rows = 2048
    for r0 in range(0, n, rows):
        block = 50 * torch.sin(x / 17.0) + 30 * torch.cos(y[r0:r0 + rows] / 11.0) + 0.05 * x + 100
        block += torch.randn(block.shape, device="cuda", generator=gen)
        block[torch.rand(block.shape, device="cuda", generator=gen) < 0.01] = NODATA
        z[0, r0:r0 + rows] = block
    yield z
    del z
    torch.cuda.empty_cache()


def window_of(z, r0, r1, c0, c1, margin):
    """Rows r0..r1 / columns c0..c1 of the OUTPUT of a stencil with `margin`: the source window."""
    return z[:, r0:r1 + 2 * margin, c0:c1 + 2 * margin].contiguous()


STENCILS = {
    "smooth": (5, lambda d: raster.Smooth.process(d, dict(smooth_mode="exact", fill=0, size=[5.0, 5.0]))),
    "moving_max": (5, lambda d: raster.MovingMax.process(d, 11)),
    "hillshade": (1, lambda d: raster.HillShade.process(
        d, dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0))),
}


@pytest.mark.parametrize("name", sorted(STENCILS))
def test_cfg3_full_size_window_invariance_and_oracle(dem32k, name):
    margin, fn = STENCILS[name]
    z = dem32k
    n = z.shape[1]
    whole = on_device(lambda: fn({"values": wrap(z), "no_data_value": NODATA}))["values"]
    assert whole.shape == (1, n - 2 * margin, n - 2 * margin)
    # view the library's result as a tensor without a host round trip
    class _View(object):
        __cuda_array_interface__ = {"shape": whole.shape, "typestr": whole.dtype.str,
                                    "data": (whole.ptr, False), "version": 2}
    whole_t = torch.as_tensor(_View(), device="cuda")
    for (r0, c0, h, w) in ((0, 0, 700, 900), (15000, 20011, 513, 1023), (n - 2 * margin - 640, n - 2 * margin - 777, 640, 777)):
        src = window_of(z, r0, r0 + h, c0, c0 + w, margin)
        part = on_device(lambda: fn({"values": wrap(src), "no_data_value": NODATA}))["values"]
        part_host = np.asarray(part)
        # a window computed on its own equals the same window of the whole raster ...
        if name == "smooth":
            # ... except where the Gaussian (radius 7) reaches beyond the 5-pixel margin: the
            # window pads with `fill` there (as the reference does), the whole raster has data
            inner = (slice(None), slice(2, h - 2), slice(2, w - 2))
            np.testing.assert_array_equal(part_host[inner], whole_t[:, r0:r0 + h, c0:c0 + w].cpu().numpy()[inner])
        else:
            np.testing.assert_array_equal(part_host, whole_t[:, r0:r0 + h, c0:c0 + w].cpu().numpy())
        # ... and the oracle (scipy.ndimage on the host copy of the source window)
        src_host = src.cpu().numpy()
        if name == "smooth":
            expected, _ = R.smooth(src_host, NODATA, (5.0, 5.0), 0, "exact")
            np.testing.assert_array_equal(part_host, expected)
        elif name == "moving_max":
            expected, _ = R.moving_max(src_host, NODATA, 11)
            np.testing.assert_array_equal(part_host, expected)
        else:
            expected, _ = R.hillshade(src_host, NODATA, (1.0, 1.0), 45.0, 315.0, 0)
            delta = np.abs(part_host.astype(int) - expected.astype(int))
            assert delta.max() <= 1 and (delta > 0).mean() <= 1e-3   # stated tolerance
    if name == "moving_max":
        # monotone: a maximum over a footprint that contains the centre is >= the centre
        centre = z[:, margin:n - margin, margin:n - margin]
        has = (centre != NODATA) & (whole_t != NODATA)
        assert bool((whole_t[has] >= centre[has]).all())


# ---- config 4: AggregateRaster over 40000 x 40000 with ~100 k polygons --------------------------------
def test_cfg4_full_size_ordering_consistency_and_sampled_oracle():
    n, g = 40000, 316
    gen = torch.Generator(device="cuda").manual_seed(4)
    r = torch.empty((1, n, n), dtype=torch.float32, device="cuda")
    rows = 4000
    for r0 in range(0, n, rows):
        block = r[0, r0:r0 + rows]
        block.uniform_(0, 100, generator=gen)
        block[torch.rand(block.shape, device="cuda", generator=gen) < 0.02] = NODATA
    rings = workloads.cfg4_rings(n, g)
    soup = utils.PolygonSoup([utils.Polygon(ring) for ring in rings]).to_device()
    bbox = (0, 0, n, n)
    rd = wrap(r)
    stats = {}
    for label, stat, q in (("count", "count", None), ("sum", "sum", None), ("mean", "mean", None),
                           ("min", "min", None), ("max", "max", None), ("median", "median", None),
                           ("p10", "percentile", 10.0), ("p90", "percentile", 90.0)):
        agg, no_cells = on_device(lambda: geometry.aggregate.aggregate_polygons(
            soup, rd, NODATA, bbox, workloads.PROJECTION, None, stat, q))
        assert no_cells == [] and agg.shape == (1, len(rings))
        stats[label] = np.array(agg[0], dtype=np.float64)
    assert np.isfinite(stats["count"]).all() and (stats["count"] > 0).all()
    # order statistics are ordered, the mean lies between the extremes and equals sum / count
    for lo, hi in (("min", "p10"), ("p10", "median"), ("median", "p90"), ("p90", "max"), ("min", "mean"), ("mean", "max")):
        assert (stats[lo] <= stats[hi]).all(), (lo, hi)
    np.testing.assert_allclose(stats["mean"], stats["sum"] / stats["count"], rtol=1e-6)
    # the oracle on 24 polygons spread over the raster, each on its own window of the raster
    rng = np.random.default_rng(5)
    for p in rng.choice(len(rings), 24, replace=False):
        ring = rings[p]
        c0, c1 = int(np.floor(ring[:, 0].min())) - 2, int(np.ceil(ring[:, 0].max())) + 2
        y_top, y_bottom = ring[:, 1].max(), ring[:, 1].min()
        r0, r1 = int(np.floor(n - y_top)) - 2, int(np.ceil(n - y_bottom)) + 2
        c0, c1, r0, r1 = max(c0, 0), min(c1, n), max(r0, 0), min(r1, n)
        frame = r[0, r0:r1, c0:c1].cpu().numpy()
        window_bbox = (c0, n - r1, c1, n - r0)
        labels = polyfill.burn_index([[ring]], window_bbox, r1 - r0, c1 - c0)
        labels = np.where(labels == 0, 0, np.iinfo(np.int32).max).astype(np.int32)
        for label, stat, q in (("count", "count", None), ("max", "max", None), ("min", "min", None),
                               ("median", "median", None), ("p90", "percentile", 90.0), ("mean", "mean", None)):
            expected, _ = R.zonal_from_labels(frame, NODATA, [(labels, [0])], 1, stat, q)
            if label == "mean":
                np.testing.assert_allclose(np.float32(stats[label][p]), expected[0], rtol=1e-6)
            else:
                assert np.float32(stats[label][p]) == expected[0], (p, label)
    del r, rd
    torch.cuda.empty_cache()


# ---- config 5: TemporalAggregate over 365 frames -----------------------------------------------------
def test_cfg5_all_365_frames_sequential_sum_and_scaling():
    from datetime import datetime, timedelta

    T, m = 365, 4096          # all 365 frames; 4096 x 4096 cells (24.5 GB) keeps the test in seconds,
    gen = torch.Generator(device="cuda").manual_seed(6)   # tools/bench_kernels.py runs 8192 x 8192
    stack = torch.empty((T, m, m), dtype=torch.float32, device="cuda")
    for t in range(T):
        stack[t].uniform_(0, 100, generator=gen)
        stack[t][torch.rand((m, m), device="cuda", generator=gen) < 0.03] = NODATA
    stack[:, 100:110, 200:260] = NODATA          # cells without any data
    times = [datetime(2000, 1, 1) + timedelta(days=i) for i in range(T)]

    def aggregate(values, statistic, dtype="f4"):
        kwargs = dict(mode="vals", start=times[-1], stop=None, frequency=None, timezone=None,
                      closed=None, label=None, dtype=dtype, statistic=statistic)
        out = on_device(lambda: raster.TemporalAggregate.process(
            kwargs, {"time": times}, {"values": wrap(values), "no_data_value": NODATA}))
        return torch.from_numpy(np.asarray(out["values"])).cuda(), out["no_data_value"]

    total, _ = aggregate(stack, "sum")
    # NumPy's axis-0 nansum accumulates sequentially in t in float32: the same loop on the device
    seq = torch.zeros((m, m), dtype=torch.float32, device="cuda")
    count = torch.zeros((m, m), dtype=torch.int32, device="cuda")
    peak = torch.full((m, m), -np.inf, dtype=torch.float32, device="cuda")
    for t in range(T):
        has = stack[t] != NODATA
        seq += torch.where(has, stack[t], torch.zeros_like(seq))
        count += has.to(torch.int32)
        peak = torch.maximum(peak, torch.where(has, stack[t], torch.full_like(peak, -np.inf)))
    assert torch.equal(total[0], seq)
    counted, _ = aggregate(stack, "count", dtype="i4")
    assert torch.equal(counted[0], count)
    highest, nodata_out = aggregate(stack, "max")
    expected_peak = torch.where(count > 0, peak, torch.full_like(peak, float(nodata_out)))
    assert torch.equal(highest[0], expected_peak)
    # scaling by a power of two commutes with the float32 sum bit for bit
    stack.mul_(0.25)
    stack[stack == np.float32(NODATA) * np.float32(0.25)] = NODATA
    scaled, _ = aggregate(stack, "sum")
    assert torch.equal(scaled[0], seq * 0.25)
    del stack
    torch.cuda.empty_cache()
