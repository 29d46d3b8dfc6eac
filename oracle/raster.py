"""NumPy/SciPy restatement of the reference ``process`` functions (oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Each function takes
and returns plain ``(values, no_data_value)`` pairs / arrays and states the
reference lines it follows (paths relative to the reference checkout).
The functions are deliberately direct: one NumPy expression per reference
statement, no attempt at speed.
"""
import numpy as np
from scipy import ndimage

# --------------------------------------------------------------------------
# helpers (utils.py:61-108, :536-547, :826-845)
# --------------------------------------------------------------------------


def dtype_max(dtype):
    d = np.dtype(dtype)
    return np.finfo(d).max.item() if d.kind == "f" else np.iinfo(d).max


def dtype_min(dtype):
    d = np.dtype(dtype)
    return np.finfo(d).min.item() if d.kind == "f" else np.iinfo(d).min


def has_data(values, nodata):
    """utils.py:61-64 (get_index)."""
    same = np.isclose(values, nodata) if values.dtype.kind == "f" else np.equal(values, nodata)
    return np.logical_not(same)


def uint_dtype(n):
    """utils.py:100-108."""
    for code in ("u1", "u2", "u4", "u8"):
        if n - 1 <= np.iinfo(code).max:
            return np.dtype(code)
    raise ValueError(n)


def int_dtype(n):
    """utils.py:91-97."""
    for code in ("i1", "i2", "i4", "i8"):
        if n - 1 <= np.iinfo(code).max and n >= np.iinfo(code).min:
            return np.dtype(code)
    raise ValueError(n)


def footprint(size):
    """utils.py:536-547."""
    s = size // 2 * 2 + 1
    o = (s - 1) // 2
    x, y = np.indices((s, s)) - o
    return (x ** 2 + y ** 2) < (s / 2) ** 2


def block_dtype(*operands):
    """raster/elemwise.py:134-144: result_type, at least int32 / float32."""
    dtype = np.result_type(*[o.dtype if isinstance(o, np.ndarray) else o for o in operands])
    if dtype == bool or np.issubdtype(dtype, np.integer):
        return np.result_type(dtype, np.int32)
    if np.issubdtype(dtype, np.floating):
        return np.result_type(dtype, np.float32)
    return dtype


# --------------------------------------------------------------------------
# element-wise (raster/elemwise.py)
# --------------------------------------------------------------------------

_UFUNCS = {
    "add": np.add, "subtract": np.subtract, "multiply": np.multiply, "divide": np.divide,
    "power": np.power, "equal": np.equal, "not_equal": np.not_equal, "greater": np.greater,
    "greater_equal": np.greater_equal, "less": np.less, "less_equal": np.less_equal,
    "logical_and": np.logical_and, "logical_or": np.logical_or, "logical_xor": np.logical_xor,
    "exp": np.exp, "log": np.log, "log10": np.log10,
}


def elementwise(name, dtype, fillvalue, *operands):
    """raster/elemwise.py:243-297.  ``operands`` are scalars or (values, nodata)
    pairs; returns (values, nodata)."""
    func = _UFUNCS[name]
    dtype = np.dtype(dtype)
    args, invalid = [], None
    for op in operands:
        if not isinstance(op, tuple):
            args.append(op)
            continue
        values, nodata = op
        args.append(values)
        if values.dtype == bool:
            continue                                            # :266-267
        hit = values == nodata                                  # :270
        invalid = hit if invalid is None else (invalid | hit)   # :271-274
    if dtype == bool:                                           # :278-284
        fill, out_nodata, kwargs = (func is np.not_equal), None, {}
    else:                                                       # :285-287
        fill, out_nodata, kwargs = fillvalue, fillvalue, {"dtype": dtype}
    with np.errstate(all="ignore"):
        result = func(*args, **kwargs)                          # :289-290
    result[~np.isfinite(result)] = fill                         # :293
    if invalid is not None:
        result[np.broadcast_to(invalid, result.shape)] = fill   # :295-296
    return result, out_nodata


def invert(values):
    """raster/elemwise.py:570-575."""
    return ~values, None


def is_data(values, nodata):
    """raster/elemwise.py:601-607."""
    return values != nodata, None


def is_nodata(values, nodata):
    """raster/elemwise.py:632-638."""
    return values == nodata, None


def fill_nodata(dtype, *rasters):
    """raster/elemwise.py:742-757: later rasters on top, no data transparent."""
    dtype = np.dtype(dtype)
    fill = dtype_max(dtype)
    out = np.full(rasters[0][0].shape, fill, dtype=dtype)
    for values, nodata in rasters:
        keep = has_data(values, nodata)
        out[keep] = values[keep]
    return out, fill


# --------------------------------------------------------------------------
# misc (raster/misc.py)
# --------------------------------------------------------------------------


def clip(values, nodata, mask_values, mask_nodata):
    """raster/misc.py:108-123."""
    if np.all(values == nodata):
        return values, nodata
    hide = ~mask_values if mask_values.dtype == bool else (mask_values == mask_nodata)
    out = values.copy()
    out[hide] = nodata
    return out, nodata


def mask(values, nodata, value):
    """raster/misc.py:199-222."""
    if isinstance(value, float):
        dtype = np.dtype("float32")
    elif value >= 0:
        dtype = uint_dtype(value)
    else:
        dtype = int_dtype(value)
    fill = 1 if value == 0 else 0
    out = np.full_like(values, fill, dtype=dtype)
    out[has_data(values, nodata)] = value
    return out, fill


def mask_below(values, nodata, value):
    """raster/misc.py:249-251."""
    out = values.copy()
    out[out < value] = nodata
    return out, nodata


def step(values, nodata, left, right, location, at):
    """raster/misc.py:314-328."""
    out = values.copy()
    missing = out == nodata
    lower, equal, higher = out < location, out == location, out > location
    out[lower] = left
    out[equal] = at
    out[higher] = right
    out[missing] = nodata
    return out, nodata


def classify(values, nodata, bins, right):
    """raster/misc.py:392-399."""
    dtype = uint_dtype(len(bins) + 2)
    fill = dtype_max(dtype)
    out = np.digitize(values, bins, right).astype(dtype)
    out[values == nodata] = fill
    return out, fill


def reclassify(values, nodata, pairs, select, dtype, fillvalue):
    """raster/misc.py:487-515."""
    source = np.asarray([p[0] for p in pairs])
    target = np.asarray([p[1] for p in pairs])
    dtype = np.dtype(dtype)
    if nodata is not None and nodata not in source:
        source = np.append(source, nodata)
        target = np.append(target, fillvalue)
    order = np.argsort(source)
    source, target = source[order], target[order]
    out = np.full(values.shape, fillvalue, dtype=dtype) if select else values.astype(dtype)
    mapped = np.isin(values.ravel(), source).reshape(values.shape)
    out[mapped] = target[np.searchsorted(source, values[mapped])]
    return out, fillvalue


# --------------------------------------------------------------------------
# spatial (raster/spatial.py)
# --------------------------------------------------------------------------


def dilate(values, nodata, dilate_values):
    """raster/spatial.py:150-155 (3-D cross: the time axis takes part)."""
    out = values.copy()
    for v in np.asarray(dilate_values, dtype=values.dtype):
        out[ndimage.binary_dilation(values == v)] = v
    return out[:, 1:-1, 1:-1], nodata


def moving_max(values, nodata, size):
    """raster/spatial.py:196-213."""
    radius = int(size // 2)
    work = values.copy()
    lowest = dtype_min(work.dtype)
    missing = work == nodata
    work[missing] = lowest
    out = ndimage.maximum_filter(work, footprint=footprint(size)[np.newaxis])
    out[(out == lowest) & missing] = nodata
    return out[:, radius:-radius, radius:-radius], nodata


def smooth(values, nodata, size_px, fill, mode):
    """raster/spatial.py:282-307; size_px = (y, x) radius in pixels."""
    work = values.copy()
    work[work == nodata] = fill
    sigma = 0, size_px[0] / 3, size_px[1] / 3
    ndimage.gaussian_filter(work, sigma, output=work, mode="constant", cval=fill)
    if mode == "exact":
        my, mx = [int(round(s)) for s in size_px]
        work = work[:, my : work.shape[1] - my, mx : work.shape[2] - mx]
    else:
        _, ny, nx = work.shape
        zy, zx = 1 - 2 * size_px[0] / ny, 1 - 2 * size_px[1] / nx
        work = ndimage.affine_transform(
            work, order=0, matrix=np.diag([1, zy, zx]), offset=[0, size_px[0], size_px[1]]
        )
    return work, nodata


def hillshade(values, nodata, resolution, altitude, azimuth, fill):
    """raster/spatial.py:364-417 (Horn gradient, zsf = 1/8)."""
    import math

    a = values.copy()
    a[a == nodata] = fill
    xres, yres = resolution
    alt, az = math.radians(altitude), math.radians(azimuth)
    zsf = 1 / 8
    n, c, s = slice(None, -2), slice(1, -1), slice(2, None)   # north/centre/south, west/centre/east
    y = np.empty(a.shape, dtype="f4")
    y[:, c, c] = (a[:, n, n] + 2 * a[:, n, c] + a[:, n, s]
                  - a[:, s, n] - 2 * a[:, s, c] - a[:, s, s]) / yres
    x = np.empty(a.shape, dtype="f4")
    x[:, c, c] = (a[:, n, n] + 2 * a[:, c, n] + a[:, s, n]
                  - a[:, n, s] - 2 * a[:, c, s] - a[:, s, s]) / xres
    with np.errstate(all="ignore"):
        xx_plus_yy = x * x + y * y
        aspect = np.arctan2(y, x)
        cang = (math.sin(alt) - math.cos(alt) * zsf * np.sqrt(xx_plus_yy) * np.sin(aspect - az)
                ) / np.sqrt(1 + zsf * zsf * xx_plus_yy)
    cang = cang[..., 1:-1, 1:-1]
    return np.where(cang <= 0, 0, 255 * cang).astype("u1"), 256


# --------------------------------------------------------------------------
# temporal (raster/temporal.py); bins are lists of frame indices
# --------------------------------------------------------------------------


def _count_valid(x, *args, **kwargs):
    return np.sum(~np.isnan(x), *args, **kwargs)


def _cumcount_valid(x, *args, **kwargs):
    return np.cumsum(~np.isnan(x), *args, **kwargs)


_TEMPORAL = {
    "sum": (np.nansum, True), "count": (_count_valid, True), "min": (np.nanmin, False),
    "max": (np.nanmax, False), "mean": (np.nanmean, False), "median": (np.nanmedian, False),
    "std": (np.nanstd, False), "var": (np.nanvar, False),
}


def statistic_dtype(dtype, statistic):
    """utils.py:826-845."""
    if statistic in ("min", "max"):
        return np.dtype(dtype)
    if statistic == "sum":
        return np.dtype(block_dtype(np.zeros(0, dtype)))
    if statistic == "count":
        return np.dtype(np.int32)
    return np.result_type(np.float32, dtype)


def temporal_aggregate(values, nodata, statistic, bins, percentile=None):
    """raster/temporal.py:729-768 with ``bins`` = frame indices per output label."""
    import warnings
    from functools import partial

    if percentile is not None:
        func, extensive = partial(np.nanpercentile, q=percentile), False
        dtype = statistic_dtype(values.dtype, "percentile")
    else:
        func, extensive = _TEMPORAL[statistic]
        dtype = statistic_dtype(values.dtype, statistic)
    fill = 0 if extensive else dtype_max(dtype)
    work = values.astype(np.result_type(np.float32, dtype))
    work[values == nodata] = np.nan
    out = np.full((len(bins),) + values.shape[1:], fill, dtype=dtype)
    for i, frames in enumerate(bins):
        if len(frames) == 0:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            agg = func(work[list(frames)], axis=0)
        bad = ~np.isfinite(agg)
        with np.errstate(all="ignore"):
            agg = agg.astype(dtype) if agg.dtype != dtype else agg
        agg[bad] = fill
        out[i] = agg
    return out, dtype_max(dtype)


def cumulative(values, nodata, statistic, bins, output_mask):
    """raster/temporal.py:966-1005: per bin nancumsum / running count; frames
    where ``output_mask`` is False are dropped."""
    func = {"sum": np.nancumsum, "count": _cumcount_valid}[statistic]
    dtype = statistic_dtype(values.dtype, statistic)
    work = values.astype(np.result_type(np.float32, dtype))
    work[values == nodata] = np.nan
    output_mask = np.asarray(output_mask, dtype=bool)
    offset = np.where(output_mask)[0][0]
    out = np.full((int(output_mask.sum()),) + values.shape[1:], 0, dtype=dtype)
    for frames in bins:
        frames = np.asarray(list(frames))
        keep = output_mask[frames]
        acc = func(work[frames], axis=0)[keep]
        bad = ~np.isfinite(acc)
        with np.errstate(all="ignore"):
            acc = acc.astype(dtype) if acc.dtype != dtype else acc
        acc[bad] = 0
        out[frames[keep] - offset] = acc
    return out, dtype_max(dtype)


# --------------------------------------------------------------------------
# zonal statistics (measurements.py, geometry/aggregate.py)
# --------------------------------------------------------------------------


def percentile(data, qval, labels, index):
    """measurements.py:93-137 for the (labels, index list) call the aggregation makes."""
    data = np.asanyarray(data)
    data, labels = np.broadcast_arrays(data, labels)
    unique_labels, labels = np.unique(labels, return_inverse=True)
    idxs = np.searchsorted(unique_labels, index)
    idxs[idxs >= unique_labels.size] = 0
    found = unique_labels[idxs] == index
    idxs[~found] = labels.max() + 1
    order = np.lexsort((data.ravel(), labels.ravel()))
    data = data.ravel()[order]
    labels = labels.ravel()[order]
    locs = np.arange(len(labels))
    lo = np.zeros(labels.max() + 2, int)
    lo[labels[::-1]] = locs[::-1]
    hi = np.zeros(labels.max() + 2, int)
    hi[labels] = locs
    lo, hi = lo[idxs], hi[idxs]
    size = hi - lo + 1
    frac = (size - 1) * (qval / 100)
    hi = lo - np.int64(-frac // 1)
    lo = lo + np.int64(frac // 1)
    part = frac % 1
    return (data[lo] + part * (data[hi] - data[lo])).tolist()


_ZONAL = {
    "sum": ndimage.sum, "count": ndimage.sum, "min": ndimage.minimum, "max": ndimage.maximum,
    "mean": ndimage.mean, "median": ndimage.median,
}


def zonal_from_labels(frame, nodata, label_sets, n_geometries, statistic, q=None, thresholds=None):
    """geometry/aggregate.py:154-203 for ONE frame.  ``label_sets`` is a list of
    (labels int32 (h, w) with int32-max where unlabelled, ids) per bucket of
    non-overlapping geometries.  Returns (agg float32 (n,), ids without cells)."""
    agg = np.full(n_geometries, np.nan, dtype="f4")
    no_cells = set()
    unlabelled = np.iinfo(np.int32).max
    if thresholds is not None:
        thresholds = np.concatenate([thresholds, np.array([np.nan], dtype=thresholds.dtype)])
    for labels, ids in label_sets:
        present = set(np.unique(labels[labels != unlabelled]).tolist())
        no_cells |= set(ids) - present
        if not present:
            continue
        active = frame != nodata
        if thresholds is not None:
            per_cell = np.take(thresholds, labels, mode="clip")
            valid = ~np.isnan(per_cell)
            active[~valid] = False
            active[valid] &= frame[valid] >= per_cell[valid]
        if not active.any():
            continue
        active_labels = labels[active]
        chosen = list(set(np.unique(active_labels)) & set(ids))
        if not chosen:
            continue
        if statistic == "percentile":
            res = percentile(frame[active], q, labels=active_labels, index=chosen)
        else:
            res = _ZONAL[statistic](1 if statistic == "count" else frame[active],
                                    labels=active_labels, index=chosen)
        agg[chosen] = res
    return agg, sorted(no_cells)


# ---- reductions over several rasters --------------------------------------------------------

_NAN_REDUCERS = {"sum": np.nansum, "min": np.nanmin, "max": np.nanmax, "product": np.nanprod}


def reduce_rasters(stack, statistic, nodata=None, dtype=None):
    """raster/reduction.py:38-119 for the statistics of the CUDA path.  ``stack`` = list of
    (values, no data value); returns (values, no data value)."""
    if dtype is None:
        dtype = stack[0][0].dtype
    if nodata is None:
        nodata = stack[0][1]
    dtype = np.dtype(dtype)
    shape = stack[0][0].shape
    out = np.full(shape, 0 if statistic in ("sum", "count") else nodata, dtype)
    if statistic in ("last", "first"):
        for values, nd in (stack if statistic == "last" else stack[::-1]):
            index = has_data_close(values, nd)
            out[index] = values[index]
    elif statistic == "count":
        for values, nd in stack:
            out += has_data_close(values, nd)
    else:
        stacked = np.full((len(stack),) + shape, np.nan, np.result_type(dtype, np.float16))
        for i, (values, nd) in enumerate(stack):
            index = has_data_close(values, nd)
            stacked[i, index] = values[index]
        some = ~np.all(np.isnan(stacked), axis=0)
        out[some] = _NAN_REDUCERS[statistic](stacked[:, some], axis=0)
    return out, nodata


def has_data_close(values, nodata):
    """utils.get_index (utils.py:61-64): np.isclose decides for floats."""
    if nodata is None:
        return np.ones(values.shape, dtype=bool)
    if values.dtype.kind == "f":
        return ~np.isclose(values, nodata)
    return values != nodata


def group_by_bands(stack, bands, dtype, shape):
    """Group._merge_vals_by_bands (raster/combine.py:371-387)."""
    fill = dtype_max(dtype)
    values = np.full(shape, fill, dtype=dtype)
    for (source, nd), (a, b) in zip(stack, bands):
        index = has_data_close(source, nd)
        values[a:b][index] = source[index]
    return values, fill


def group_by_time(stack, times, dtype, start, stop):
    """Group._merge_vals_by_time (raster/combine.py:316-343)."""
    instants = sorted(set(t for ts in times for t in ts))
    frame_of = {t: k for k, t in enumerate(instants)}
    fill = dtype_max(dtype)
    values = np.full((len(instants),) + stack[0][0].shape[1:], fill, dtype=dtype)
    for (source, nd), ts in zip(stack, times):
        for i, t in enumerate(ts):
            index = has_data_close(source[i], nd)
            values[frame_of[t]][index] = source[i][index]
    if stop is None and len(instants) > 1:
        k = len(instants) - 1 if start is None else min(range(len(instants)), key=lambda i: abs(instants[i] - start))
        values = values[k:k + 1]
    return values, fill


def place_warp(values, nodata, kwargs):
    """Place.process, mode "warp" (raster/spatial.py:657-731): the source is shifted onto every
    coordinate and the copies are merged with reduce_rasters."""
    size_x, size_y = kwargs["cellsize"]
    anchor, src_bbox = kwargs["anchor"], kwargs["src_bbox"]
    anchor_px = ((anchor[0] - src_bbox[0]) / size_x, (anchor[1] - src_bbox[1]) / size_y)
    x1, y1, x2, y2 = kwargs["dst_bbox"]
    dst_h, dst_w = round((y2 - y1) / size_y), round((x2 - x1) / size_x)
    depth, src_h, src_w = values.shape
    shape = (depth, dst_h, dst_w)
    k, j, i = np.where(has_data_close(values, nodata))
    stack = []
    for x, y in kwargs["coordinates"]:
        if i.size == 0:
            break
        di = round((x - x1) / size_x - anchor_px[0])
        dj = dst_h - src_h - round((y - y1) / size_y - anchor_px[1])
        if di <= -src_w or di >= dst_w or dj <= -src_h or dj >= dst_h:
            continue
        i_s, j_s = i + di, j + dj
        m = (i_s >= 0) & (j_s >= 0) & (i_s < dst_w) & (j_s < dst_h)
        if not m.any():
            continue
        placed = np.full(shape, nodata, values.dtype)
        placed[k[m], j_s[m], i_s[m]] = values[k[m], j[m], i[m]]
        stack.append((placed, nodata))
    if not stack:
        return np.full(shape, nodata, values.dtype), nodata
    return reduce_rasters(stack, kwargs["statistic"])
