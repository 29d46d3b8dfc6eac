"""Replay the golden cases of tests/golden/*.json (outputs of the reference's own
functions, see tests/golden/make_golden.py) through the oracle or the CUDA blocks."""
import json
import os
from datetime import datetime

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAMILIES = ["elemwise", "misc", "spatial", "temporal", "zonal", "reduce"]

# ops whose NumPy/libm result cannot be reproduced bit for bit: stated tolerances
TRANSCENDENTAL = {"power", "exp", "log", "log10"}
ULP_TOLERANT_TEMPORAL = set()  # every temporal statistic is reproduced bit for bit


def load(family):
    with open(os.path.join(GOLDEN, family + ".json")) as f:
        cases = json.load(f)
    arrays = np.load(os.path.join(GOLDEN, family + ".npz"))
    return cases, arrays


def all_cases():
    out = []
    for family in FAMILIES:
        cases, arrays = load(family)
        out.extend((family, c, arrays) for c in cases)
    return out


def inputs_of(case, arrays):
    return [(arrays[i["key"]], i["nodata"]) for i in case["inputs"]]


def expected_of(case, arrays):
    if case["output"] is None:
        return None
    return arrays[case["output"]["key"]], case["output"]["nodata"]


def _dt(s):
    return None if s is None else datetime.fromisoformat(s)


def _temporal_kwargs(args):
    kwargs = dict(args["kwargs"])
    kwargs["start"], kwargs["stop"] = _dt(kwargs["start"]), _dt(kwargs["stop"])
    return kwargs, [_dt(t) for t in args["times"]]


# ------------------------------------------------------------------ oracle ----
def run_oracle(case, arrays):
    from oracle import raster as R

    ins = inputs_of(case, arrays)
    fam, op, args = case["family"], case["op"], case["args"]
    if fam == "math":
        it = iter(ins)
        operands = [next(it) if o == "in" else o for o in args["operands"]]
        return R.elementwise(op, args["kwargs"]["dtype"], args["kwargs"]["fillvalue"], *operands)
    if fam == "invert":
        return R.invert(ins[0][0])
    if fam == "isdata":
        return R.is_data(*ins[0])
    if fam == "isnodata":
        return R.is_nodata(*ins[0])
    if fam == "fillnodata":
        return R.fill_nodata(args["kwargs"]["dtype"], *ins)
    if fam == "clip":
        return R.clip(ins[0][0], ins[0][1], ins[1][0], ins[1][1])
    if fam == "mask":
        return R.mask(ins[0][0], ins[0][1], args["value"])
    if fam == "maskbelow":
        return R.mask_below(ins[0][0], ins[0][1], args["value"])
    if fam == "step":
        return R.step(ins[0][0], ins[0][1], *args["args"])
    if fam == "classify":
        return R.classify(ins[0][0], ins[0][1], args["bins"], args["right"])
    if fam == "reclassify":
        k = args["kwargs"]
        return R.reclassify(ins[0][0], ins[0][1], k["data"], k["select"], k["dtype"], k["fillvalue"])
    if fam == "dilate":
        return R.dilate(ins[0][0], ins[0][1], args["values"])
    if fam == "movingmax":
        return R.moving_max(ins[0][0], ins[0][1], args["size"])
    if fam == "smooth":
        k = args["kwargs"]
        return R.smooth(ins[0][0], ins[0][1], k["size"], k["fill"], k["smooth_mode"])
    if fam == "hillshade":
        k = args["kwargs"]
        return R.hillshade(ins[0][0], ins[0][1], k["resolution"], k["altitude"], k["azimuth"], k["fill"])
    if fam in ("temporal_aggregate", "cumulative"):
        kwargs, times = _temporal_kwargs(args)
        values, nodata = ins[0]
        n = len(times)
        if fam == "temporal_aggregate":
            stat = kwargs["statistic"]
            q = float(stat[1:]) if stat.startswith("p") else None
            bins = [range(n)] if kwargs["frequency"] is None else [range(0, 4), range(4, 8), range(8, min(n, 12))]
            if kwargs["frequency"] is not None:
                bins = bins[: 3]
            return R.temporal_aggregate(values, nodata, "percentile" if q is not None else stat, bins, q)
        bins = [list(range(n))] if kwargs["frequency"] is None else [[0], [1, 2, 3, 4], [5, 6, 7, 8], [9]]
        mask = np.array([kwargs["start"] <= t <= kwargs["stop"] for t in times])
        return R.cumulative(values, nodata, kwargs["statistic"], bins, mask)
    if fam == "percentile":
        frame, labels = ins[0][0], ins[1][0]
        active = frame != np.finfo("f4").max
        return np.asarray(R.percentile(frame[active], args["q"], labels[active], args["index"])), None
    if fam == "labelled":
        frame, labels = ins[0][0], ins[1][0]
        active = frame != np.finfo("f4").max
        return np.asarray(R._ZONAL[op](frame[active], labels=labels[active], index=args["index"])), None
    if fam == "reduce":
        return R.reduce_rasters(ins, op, args["no_data_value"], args["dtype"])
    if fam == "group_bands":
        return R.group_by_bands(ins, args["bands"], args["dtype"], tuple(args["shape"]))
    if fam == "group_time":
        return R.group_by_time(ins, args["times"], args["dtype"], args["start"], args["stop"])
    if fam == "place_warp":
        return R.place_warp(ins[0][0], ins[0][1], args)
    raise KeyError(fam)


# ------------------------------------------------------------------ product ---
def run_product(case, arrays):
    from dask_geomodeling_b200 import raster

    ins = [{"values": v, "no_data_value": nd} for v, nd in inputs_of(case, arrays)]
    fam, op, args = case["family"], case["op"], case["args"]
    blocks = {
        "add": raster.Add, "subtract": raster.Subtract, "multiply": raster.Multiply,
        "divide": raster.Divide, "power": raster.Power, "equal": raster.Equal,
        "not_equal": raster.NotEqual, "greater": raster.Greater,
        "greater_equal": raster.GreaterEqual, "less": raster.Less, "less_equal": raster.LessEqual,
        "exp": raster.Exp, "log": raster.Log, "log10": raster.Log10, "logical_and": raster.And,
        "logical_or": raster.Or, "logical_xor": raster.Xor,
    }
    if fam == "math":
        it = iter(ins)
        operands = [next(it) if o == "in" else o for o in args["operands"]]
        res = blocks[op].process(args["kwargs"], *operands)
    elif fam == "invert":
        res = raster.Invert.process(ins[0])
    elif fam == "isdata":
        res = raster.IsData.process(ins[0])
    elif fam == "isnodata":
        res = raster.IsNoData.process(ins[0])
    elif fam == "fillnodata":
        res = raster.FillNoData.process(args["kwargs"], *ins)
    elif fam == "clip":
        res = raster.Clip.process(ins[0], ins[1])
    elif fam == "mask":
        res = raster.Mask.process(ins[0], args["value"])
    elif fam == "maskbelow":
        res = raster.MaskBelow.process(ins[0], args["value"])
    elif fam == "step":
        res = raster.Step.process(ins[0], *args["args"])
    elif fam == "classify":
        res = raster.Classify.process(ins[0], args["bins"], args["right"])
    elif fam == "reclassify":
        res = raster.Reclassify.process(ins[0], args["kwargs"])
    elif fam == "dilate":
        res = raster.Dilate.process(ins[0], args["values"])
    elif fam == "movingmax":
        res = raster.MovingMax.process(ins[0], args["size"])
    elif fam == "smooth":
        res = raster.Smooth.process(ins[0], args["kwargs"])
    elif fam == "hillshade":
        res = raster.HillShade.process(ins[0], args["kwargs"])
    elif fam == "temporal_aggregate":
        kwargs, times = _temporal_kwargs(args)
        res = raster.TemporalAggregate.process(kwargs, {"time": times}, ins[0])
    elif fam == "cumulative":
        kwargs, times = _temporal_kwargs(args)
        res = raster.Cumulative.process(kwargs, {"time": times}, ins[0])
    elif fam == "reduce":
        from dask_geomodeling_b200.raster.reduction import reduce_rasters

        res = reduce_rasters(ins, op, args["no_data_value"], args["dtype"])
    elif fam == "group_bands":
        res = raster.Group._merge_vals_by_bands(ins, [tuple(b) for b in args["bands"]], np.dtype(args["dtype"]),
                                                tuple(args["shape"]))
    elif fam == "place_warp":
        res = raster.Place.process(args, ins[0])
    elif fam == "group_time":
        res = raster.Group._merge_vals_by_time(
            ins, [{"time": t} for t in args["times"]],
            {"dtype": np.dtype(args["dtype"]), "start": args["start"], "stop": args["stop"]})
    else:
        raise KeyError(fam)
    return np.asarray(res["values"]), res["no_data_value"]


def compare(case, got, expected):
    """Bit-exact unless the op is listed with a stated tolerance."""
    values, nodata = got
    e_values, e_nodata = expected
    fam, op = case["family"], case["op"]
    assert values.dtype == e_values.dtype, (fam, op, values.dtype, e_values.dtype)
    assert values.shape == e_values.shape
    if e_nodata is None:
        assert nodata is None
    else:
        assert nodata == e_nodata
    if fam == "math" and op in TRANSCENDENTAL and e_values.dtype.kind == "f":
        # few-ulp differences between libm/SIMD (NumPy) and CUDA math functions
        np.testing.assert_array_equal(values == e_nodata, e_values == e_nodata)
        ok = e_values != e_nodata
        np.testing.assert_allclose(values[ok], e_values[ok], rtol=1e-6 if e_values.dtype == np.float32 else 1e-14)
    elif fam == "hillshade":
        delta = np.abs(values.astype(int) - e_values.astype(int))
        assert delta.max() <= 1 and (delta > 0).mean() <= 1e-3
    elif fam == "temporal_aggregate" and op in ULP_TOLERANT_TEMPORAL:
        np.testing.assert_allclose(values, e_values, rtol=3e-7 if e_values.dtype == np.float32 else 1e-15)
    else:
        np.testing.assert_array_equal(values, e_values)
