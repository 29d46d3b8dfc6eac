"""Generate the golden vectors in this directory by running the REFERENCE's own
``process`` functions (imported from /root/reference behind stubbed third-party
packages, see oracle/refharness.py) on small seeded inputs.

    python tests/golden/make_golden.py

Only runs in the build container (the reference checkout does not travel to
the GPU box); the resulting ``*.npz`` files are committed.  Every case stores
the call arguments, the inputs and the reference output, so that both the
oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_golden_gpu.py) can be checked against the reference itself.
"""
import json
import os
import sys
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refharness  # noqa: E402

SHAPE = (2, 23, 31)


def dmax(dtype):
    d = np.dtype(dtype)
    return np.finfo(d).max.item() if d.kind == "f" else int(np.iinfo(d).max)


def raster(dtype, seed, lo=0, hi=100, nodata_fraction=0.15, shape=SHAPE):
    rng = np.random.default_rng(seed)
    dtype = np.dtype(dtype)
    if dtype == bool:
        return rng.random(shape) < 0.5, None
    nodata = dmax(dtype)
    values = (rng.uniform(lo, hi, shape) if dtype.kind == "f" else rng.integers(lo, hi, shape)).astype(dtype)
    values[rng.random(shape) < nodata_fraction] = nodata
    return values, nodata


def payload(pair):
    return {"values": pair[0], "no_data_value": pair[1]}


class Recorder(object):
    def __init__(self):
        self.arrays, self.cases = {}, []

    def add(self, family, op, args, inputs, result):
        """args: JSON-able call parameters; inputs: list of (values, nodata)."""
        idx = len(self.cases)
        case = {"family": family, "op": op, "args": args, "inputs": [], "id": idx}
        for i, (values, nodata) in enumerate(inputs):
            key = "c{}_in{}".format(idx, i)
            self.arrays[key] = values
            case["inputs"].append({"key": key, "nodata": nodata})
        if result is None:
            case["output"] = None
        else:
            key = "c{}_out".format(idx)
            self.arrays[key] = np.asarray(result["values"])
            nd = result["no_data_value"]
            case["output"] = {"key": key, "nodata": nd if nd is None else
                              (float(nd) if isinstance(nd, (float, np.floating)) else int(nd))}
        self.cases.append(case)

    def save(self, name):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **self.arrays)
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(self.cases, f, indent=0, default=lambda o: o.item() if hasattr(o, "item") else str(o))
        print(name, len(self.cases), "cases")


def block_dtype(*ops):
    dtype = np.result_type(*[o[0].dtype if isinstance(o, tuple) else o for o in ops])
    if dtype == bool or np.issubdtype(dtype, np.integer):
        return np.result_type(dtype, np.int32)
    return np.result_type(dtype, np.float32)


def elementwise(ns):
    rec = Recorder()
    E = ns.elemwise
    math_blocks = {"add": E.Add, "subtract": E.Subtract, "multiply": E.Multiply,
                   "divide": E.Divide, "power": E.Power}
    cmp_blocks = {"equal": E.Equal, "not_equal": E.NotEqual, "greater": E.Greater,
                  "greater_equal": E.GreaterEqual, "less": E.Less, "less_equal": E.LessEqual}
    seed = 100
    for dtypes in [("f4", "f4"), ("u1", "i2"), ("i4", "f4"), ("f4", "f8"), ("u1", "u1")]:
        a, b = raster(dtypes[0], seed, hi=12), raster(dtypes[1], seed + 1, hi=6)
        seed += 2
        for name, block in list(math_blocks.items()) + list(cmp_blocks.items()):
            if name in cmp_blocks:
                dtype, fill = np.dtype(bool), None
            elif name == "divide":
                dtype = np.result_type(np.float32, a[0].dtype, b[0].dtype)
                fill = dmax(dtype)
            else:
                dtype = block_dtype(a, b)
                fill = dmax(dtype)
            kwargs = {"dtype": dtype.name, "fillvalue": fill}
            rec.add("math", name, {"kwargs": kwargs, "operands": ["in", "in"]}, [a, b],
                    block.process(kwargs, payload(a), payload(b)))
    for dtype_in in ["f4", "u1", "i2"]:
        a = raster(dtype_in, seed, hi=9)
        seed += 1
        for scalar in (3, 2.5):
            for first in (False, True):
                for name, block in list(math_blocks.items()) + list(cmp_blocks.items()):
                    ops = (scalar, a) if first else (a, scalar)
                    if name in cmp_blocks:
                        dtype, fill = np.dtype(bool), None
                    elif name == "divide":
                        dtype = np.result_type(np.float32, *[o[0].dtype if isinstance(o, tuple) else o for o in ops])
                        fill = dmax(dtype)
                    else:
                        dtype = block_dtype(*ops)
                        fill = dmax(dtype)
                    kwargs = {"dtype": dtype.name, "fillvalue": fill}
                    spec = [("in" if isinstance(o, tuple) else o) for o in ops]
                    rec.add("math", name, {"kwargs": kwargs, "operands": spec}, [a],
                            block.process(kwargs, *[payload(o) if isinstance(o, tuple) else o for o in ops]))
    for name, block in [("exp", E.Exp), ("log", E.Log), ("log10", E.Log10)]:
        for dtype_in in ("f4", "i2"):
            a = raster(dtype_in, seed, lo=-3, hi=95)
            seed += 1
            dtype = np.result_type(np.float32, a[0].dtype)
            kwargs = {"dtype": dtype.name, "fillvalue": dmax(dtype)}
            rec.add("math", name, {"kwargs": kwargs, "operands": ["in"]}, [a], block.process(kwargs, payload(a)))
    for name, block in [("logical_and", E.And), ("logical_or", E.Or), ("logical_xor", E.Xor)]:
        a, b = raster(bool, seed), raster(bool, seed + 1)
        seed += 2
        kwargs = {"dtype": "bool", "fillvalue": None}
        rec.add("math", name, {"kwargs": kwargs, "operands": ["in", "in"]}, [a, b],
                block.process(kwargs, payload(a), payload(b)))
        rec.add("math", name, {"kwargs": kwargs, "operands": ["in", True]}, [a],
                block.process(kwargs, payload(a), True))
    a = raster(bool, seed)
    rec.add("invert", "invert", {}, [a], E.Invert.process(payload(a)))
    for dtype_in in ("u1", "f4", "i4"):
        a = raster(dtype_in, seed)
        seed += 1
        rec.add("isdata", "isdata", {}, [a], E.IsData.process(payload(a)))
        rec.add("isnodata", "isnodata", {}, [a], E.IsNoData.process(payload(a)))
    for dtypes in [("f4", "f4", "f4"), ("u1", "i2", "u1")]:
        rs = [raster(d, seed + i, nodata_fraction=0.5) for i, d in enumerate(dtypes)]
        seed += 3
        dtype = block_dtype(*rs)
        kwargs = {"dtype": dtype.name, "fillvalue": dmax(dtype)}
        rec.add("fillnodata", "fillnodata", {"kwargs": kwargs}, rs,
                E.FillNoData.process(kwargs, *[payload(r) for r in rs]))
    rec.save("elemwise")


def misc(ns):
    rec = Recorder()
    M = ns.misc
    seed = 200
    for dtype in ("f4", "u1", "i2"):
        a = raster(dtype, seed)
        for mk in ("bool", "u1", "f4"):
            m = raster(mk, seed + 1, nodata_fraction=0.4)
            rec.add("clip", "clip", {}, [a, m], M.Clip.process(payload(a), payload(m)))
        for value in (1, 0, 7.5, -3, 300):
            rec.add("mask", "mask", {"value": value}, [a], M.Mask.process(payload(a), value))
        for value in (50, 33.3):
            rec.add("maskbelow", "maskbelow", {"value": value}, [a], M.MaskBelow.process(payload(a), value))
        for args in [(0, 1, 50, 0.5), (10, 20, 33.3, 15.0)]:
            rec.add("step", "step", {"args": list(args)}, [a], M.Step.process(payload(a), *args))
        for bins in ([10, 20, 50], [0.5, 33.3, 66.6, 99.9]):
            for right in (False, True):
                rec.add("classify", "classify", {"bins": bins, "right": right}, [a],
                        M.Classify.process(payload(a), bins, right))
        seed += 2
    for dtype in ("u1", "i2", "i4"):
        a = raster(dtype, seed)
        seed += 1
        for pairs in ([[k, 10 * k] for k in range(0, 100, 2)], [[k, k / 4] for k in range(0, 100, 3)],
                      [[5, 1], [90, 2], [dmax(dtype), 3]]):
            for select in (False, True):
                target_dtype = np.asarray([p[1] for p in pairs]).dtype
                kwargs = {"dtype": target_dtype.str, "fillvalue": dmax(target_dtype), "data": pairs,
                          "select": select}
                rec.add("reclassify", "reclassify", {"kwargs": kwargs}, [a],
                        M.Reclassify.process(payload(a), kwargs))
    rec.save("misc")


def dem(shape, seed, dtype="f4", nodata_fraction=0.02):
    rng = np.random.default_rng(seed)
    t, h, w = shape
    y, x = np.mgrid[0:h, 0:w]
    base = 50 * np.sin(x / 17.0) + 30 * np.cos(y / 11.0) + 0.05 * x + rng.normal(0, 1, (t, h, w))
    values = (base + 100).astype(dtype)
    nodata = dmax(dtype)
    values[rng.random(shape) < nodata_fraction] = nodata
    return values, nodata


def spatial(ns):
    rec = Recorder()
    S = ns.spatial
    rng = np.random.default_rng(300)
    for dtype in ("u1", "i4"):
        values = rng.integers(0, 6, (3, 20, 27)).astype(dtype)
        a = (values, dmax(dtype))
        rec.add("dilate", "dilate", {"values": [3, 1, 5]}, [a], S.Dilate.process(payload(a), [3, 1, 5]))
    for dtype in ("u1", "f4"):
        a = dem((2, 40, 53), 301, dtype=dtype, nodata_fraction=0.3)
        a[0][0, 10:30, 10:40] = a[1]
        for size in (3, 5, 11):
            rec.add("movingmax", "movingmax", {"size": size}, [a], S.MovingMax.process(payload(a), size))
    for dtype in ("f4", "i2"):
        a = dem((2, 50, 64), 302, dtype=dtype)
        for size_px, mode, fill in [((5.0, 5.0), "exact", 0), ((2.0, 3.4), "exact", 0), ((3.7, 4.2), "zoom", 7.5)]:
            kwargs = dict(smooth_mode=mode, fill=fill, size=list(size_px))
            rec.add("smooth", "smooth", {"kwargs": kwargs}, [a], S.Smooth.process(payload(a), kwargs))
    for dtype in ("f4", "f8"):
        a = dem((2, 45, 61), 303, dtype=dtype)
        for angles in [(45.0, 315.0), (30.0, 100.0)]:
            kwargs = dict(resolution=(0.5, 0.5), altitude=angles[0], azimuth=angles[1], fill=0)
            rec.add("hillshade", "hillshade", {"kwargs": kwargs}, [a], S.HillShade.process(payload(a), kwargs))
    rec.save("spatial")


def temporal(ns):
    rec = Recorder()
    T = ns.temporal
    ts = [datetime(2000, 1, 1) + timedelta(hours=i) for i in range(10)]
    iso = [t.isoformat() for t in ts]
    for dtype in ("f4", "u1", "i4"):
        a = raster(dtype, 400, nodata_fraction=0.3, shape=(10, 9, 14))
        a[0][:, 0, 0] = a[1]
        for statistic in ("sum", "count", "min", "max", "mean", "median", "std", "var", "p90"):
            out_dtype = np.dtype(ns.utils.dtype_for_statistic(a[0].dtype, statistic))
            kwargs = dict(mode="vals", start=ts[-1], stop=None, frequency=None, timezone=None,
                          closed=None, label=None, dtype=out_dtype.str, statistic=statistic)
            jk = dict(kwargs, start=iso[-1])
            rec.add("temporal_aggregate", statistic, {"kwargs": jk, "times": iso}, [a],
                    T.TemporalAggregate.process(kwargs, {"time": ts}, payload(a)))
        for statistic in ("sum", "max"):
            out_dtype = np.dtype(ns.utils.dtype_for_statistic(a[0].dtype, statistic))
            kwargs = dict(mode="vals", start=ts[0], stop=ts[8], frequency="4h", timezone="UTC",
                          closed="left", label="left", dtype=out_dtype.str, statistic=statistic)
            jk = dict(kwargs, start=iso[0], stop=iso[8])
            rec.add("temporal_aggregate", statistic, {"kwargs": jk, "times": iso}, [a],
                    T.TemporalAggregate.process(kwargs, {"time": ts}, payload(a)))
        for statistic in ("sum", "count"):
            for frequency in (None, "4h"):
                out_dtype = np.dtype(ns.utils.dtype_for_statistic(a[0].dtype, statistic))
                kwargs = dict(mode="vals", start=ts[2], stop=ts[8], frequency=frequency,
                              timezone=None if frequency is None else "UTC", closed="right",
                              label="right", dtype=out_dtype.str, statistic=statistic)
                jk = dict(kwargs, start=iso[2], stop=iso[8])
                rec.add("cumulative", statistic, {"kwargs": jk, "times": iso}, [a],
                        T.Cumulative.process(kwargs, {"time": ts}, payload(a)))
    rec.save("temporal")


def zonal(ns):
    """measurements.percentile and the scipy.ndimage labelled statistics as
    aggregate_polygons calls them (labels with int32-max where unlabelled)."""
    rec = Recorder()
    rng = np.random.default_rng(500)
    frame = rng.uniform(0, 100, (40, 50)).astype("f4")
    labels = np.full((40, 50), np.iinfo(np.int32).max, dtype=np.int32)
    for k in range(12):
        i, j = (k // 4) * 13, (k % 4) * 12
        labels[i : i + 11, j : j + 10] = k
    frame[rng.random(frame.shape) < 0.1] = np.finfo("f4").max
    active = frame != np.finfo("f4").max
    index = sorted(set(np.unique(labels[active]).tolist()) - {np.iinfo(np.int32).max})
    for q in (90.0, 50.0, 12.5, 0.0, 100.0):
        res = ns.measurements.percentile(frame[active], q, labels=labels[active], index=index)
        rec.add("percentile", "percentile", {"q": q, "index": index}, [(frame, None), (labels, None)],
                {"values": np.asarray(res, dtype=np.float64), "no_data_value": None})
    from scipy import ndimage
    for name, func in [("sum", ndimage.sum), ("mean", ndimage.mean), ("min", ndimage.minimum),
                       ("max", ndimage.maximum), ("median", ndimage.median)]:
        res = func(frame[active], labels=labels[active], index=index)
        rec.add("labelled", name, {"index": index}, [(frame, None), (labels, None)],
                {"values": np.asarray(res), "no_data_value": None})
    rec.save("zonal")


def reductions(ns):
    """reduce_rasters (raster/reduction.py:38-119) as Max and Place call it, and Group's two
    merges (raster/combine.py:316-343, :371-387)."""
    rec = Recorder()
    seed = 700
    for dtype in ("f4", "f8", "u1", "i2", "i4"):
        for statistic in ("last", "first", "count", "max", "min", "sum", "product"):
            if statistic in ("sum", "product") and dtype not in ("f4", "f8"):
                continue
            seed += 1
            stack = [raster(dtype, seed * 10 + k, lo=1, hi=9 if statistic == "product" else 100,
                            nodata_fraction=0.4) for k in range(3)]
            # a second 'no data' value in the middle raster (sources need not agree on it)
            values, nodata = stack[1]
            other = np.dtype(dtype).type(7)
            values = values.copy()
            values[values == nodata] = other
            stack[1] = (values, other.item())
            res = ns.reduction.reduce_rasters([payload(p) for p in stack], statistic, dmax(dtype), dtype)
            rec.add("reduce", statistic, {"dtype": dtype, "no_data_value": dmax(dtype)}, stack, res)
    # Max over mixed dtypes: output dtype = result_type, no data value of the block
    stack = [raster("u1", 801, nodata_fraction=0.5), raster("f4", 802, nodata_fraction=0.5)]
    res = ns.reduction.reduce_rasters([payload(p) for p in stack], "max", dmax("f4"), "float32")
    rec.add("reduce", "max", {"dtype": "float32", "no_data_value": dmax("f4")}, stack, res)
    # defaults: dtype and no data value of the first raster
    stack = [raster("i2", 803, nodata_fraction=0.5), raster("i2", 804, nodata_fraction=0.5)]
    res = ns.reduction.reduce_rasters([payload(p) for p in stack], "last")
    rec.add("reduce", "last", {"dtype": None, "no_data_value": None}, stack, res)
    # Group, equidistant sources: frames [0, 2) from the first, [1, 3) from the second, ...
    for dtype in ("f4", "u1"):
        a, b, c = (raster(dtype, 810 + k, nodata_fraction=0.5, shape=(2, 23, 31)) for k in range(3))
        bands = [(0, 2), (1, 3), (2, 4)]
        res = ns.combine.Group._merge_vals_by_bands(
            [payload(a), payload(b), payload(c)], bands, np.dtype(dtype), (4, 23, 31))
        rec.add("group_bands", "group", {"bands": bands, "dtype": dtype, "shape": [4, 23, 31]}, [a, b, c], res)
    # Group, sources with their own time stamps (integers stand in for datetimes)
    a, b = raster("f4", 820, nodata_fraction=0.5, shape=(3, 23, 31)), raster("f4", 821, nodata_fraction=0.5)
    times = [[0, 2, 5], [2, 3]]
    for start, stop in ((0, 5), (3, None), (None, None)):
        res = ns.combine.Group._merge_vals_by_time(
            [payload(a), payload(b)], [{"time": t} for t in times],
            {"dtype": np.dtype("f4"), "start": start, "stop": stop})
        rec.add("group_time", "group", {"times": times, "dtype": "f4", "start": start, "stop": stop}, [a, b], res)
    # Place, warp mode (raster/spatial.py:657-731): one source shifted onto every coordinate
    src = raster("f4", 830, nodata_fraction=0.3, shape=(2, 6, 8))
    for statistic in ("last", "first", "max", "count", "sum"):
        kwargs = {"mode": "warp", "anchor": (12.0, 22.0), "src_bbox": (10.0, 20.0, 18.0, 26.0),
                  "dst_bbox": (0.0, 0.0, 30.0, 20.0), "cellsize": (1.0, 1.0), "statistic": statistic,
                  "coordinates": [(5.0, 5.0), (12.0, 10.0), (29.0, 19.0), (100.0, 100.0), (14.0, 6.0), (1.0, 1.0)]}
        res = ns.spatial.Place.process(kwargs, payload(src))
        rec.add("place_warp", statistic, kwargs, [src], res)
    kwargs = {"mode": "warp", "anchor": (12.0, 22.0), "src_bbox": (10.0, 20.0, 18.0, 26.0),
              "dst_bbox": (0.0, 0.0, 30.0, 20.0), "cellsize": (1.0, 1.0), "statistic": "last",
              "coordinates": [(500.0, 5.0)]}
    rec.add("place_warp", "last", kwargs, [src], ns.spatial.Place.process(kwargs, payload(src)))
    rec.save("reduce")


if __name__ == "__main__":
    ns = refharness.load()
    only = set(sys.argv[1:])
    for name, build in (("elemwise", elementwise), ("misc", misc), ("spatial", spatial),
                        ("temporal", temporal), ("zonal", zonal), ("reduce", reductions)):
        if not only or name in only:
            build(ns)
