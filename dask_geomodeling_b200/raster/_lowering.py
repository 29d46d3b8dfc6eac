"""``process`` functions of the element-wise / misc blocks and their lowering
to expression nodes.

Every function below keeps the reference's calling convention and its
handling of ``None`` / time / meta payloads, builds a one-node expression
and evaluates it on the GPU.  Each also carries ``_gm_lower`` so that the
graph optimiser (core/fusion.py) can splice the same node into a larger
expression instead of launching it on its own.
"""
import numpy as np

from .. import _state
from . import _program
from ._program import Leaf, Node


def _is_values(data):
    return isinstance(data, dict) and "values" in data


def run_single(build, payloads):
    """Evaluate ``build(operands)`` where payload dicts become leaves."""
    leaves, operands = [], []
    for item in payloads:
        if _is_values(item):
            operands.append(Leaf(len(leaves)))
            leaves.append((item["values"], item.get("no_data_value")))
        elif isinstance(item, np.ndarray):
            operands.append(Leaf(len(leaves)))
            leaves.append((item, None))
        else:
            operands.append(item)
    node = build(operands)
    (values, dtype, nodata), = _program.evaluate([node], leaves, _state.keep_on_device())
    return {"values": values, "no_data_value": nodata}


# -- math / comparison / logic (raster/elemwise.py:235-299) -----------------------


def math_process(name):
    def lower(args):
        kwargs = args[0]
        return Node(name, args[1:], dtype=kwargs["dtype"], fillvalue=kwargs["fillvalue"])

    def math_process_func(process_kwargs, *args):
        if not args:
            return None
        for data in args:
            if data is None:
                return None
            if not isinstance(data, dict):
                continue
            if "time" in data or "meta" in data:
                return data  # time / meta requests are answered by the first source
            if "values" not in data:
                raise TypeError("Cannot apply math function to value {}".format(data))
        result = run_single(lambda ops: lower([process_kwargs] + ops), args)
        if np.dtype(process_kwargs["dtype"]) == bool:
            result["no_data_value"] = None
        return result

    math_process_func.__name__ = name + "_process"
    math_process_func._gm_lower = lower
    math_process_func._gm_array_operands = True   # ndarray literals are raster operands
    return math_process_func


# -- single raster ops (raster/elemwise.py:570-638) -----------------------------------


def _unary(op):
    def lower(args):
        return Node(op, args[:1])

    return lower


def invert_process(data):
    if not _is_values(data):
        return data
    return run_single(_unary("invert"), [data])


def isdata_process(data):
    if not _is_values(data):
        return data
    return run_single(_unary("isdata"), [data])


def isnodata_process(data):
    if not _is_values(data):
        return data
    return run_single(_unary("isnodata"), [data])


invert_process._gm_lower = _unary("invert")
isdata_process._gm_lower = _unary("isdata")
isnodata_process._gm_lower = _unary("isnodata")


# -- FillNoData (raster/elemwise.py:726-757) ----------------------------------------------


def _lower_fillnodata(args):
    return Node("fillnodata", args[1:], dtype=args[0]["dtype"])


def fillnodata_process(process_kwargs, *args):
    rasters = []
    for data in args:
        if data is None:
            continue
        if "time" in data or "meta" in data:
            return data
        if "values" in data and "no_data_value" in data:
            rasters.append(data)
    if not rasters:
        return None
    return run_single(lambda ops: _lower_fillnodata([process_kwargs] + ops), rasters)


fillnodata_process._gm_lower = _lower_fillnodata
fillnodata_process._gm_skip_none = True


# -- misc blocks (raster/misc.py) -------------------------------------------------------------


def _lower_clip(args):
    return Node("clip", args[:2])


def clip_process(data, source_data):
    """Clip ``data`` to the cells where ``source_data`` has data (misc.py:98-123)."""
    if data is None:
        return None
    if "values" not in data:
        return data
    if source_data is None:
        # reference order: an all-'no data' raster is returned as is, before the
        # missing mask turns the result into None (misc.py:108-113)
        nodata_everywhere = run_single(_unary("isnodata"), [data])["values"]
        return data if bool(np.asarray(nodata_everywhere).all()) else None
    return run_single(_lower_clip, [data, source_data])


clip_process._gm_lower = _lower_clip


def _lower_mask(args):
    return Node("mask", args[:1], value=args[1])


def mask_process(data, value):
    if not _is_values(data):
        return data
    return run_single(lambda ops: _lower_mask(ops + [value]), [data])


mask_process._gm_lower = _lower_mask


def _lower_maskbelow(args):
    return Node("maskbelow", args[:1], value=args[1])


def maskbelow_process(data, value):
    if not _is_values(data):
        return data
    return run_single(lambda ops: _lower_maskbelow(ops + [value]), [data])


maskbelow_process._gm_lower = _lower_maskbelow


def _lower_step(args):
    data, left, right, location, at = args[:5]
    return Node("step", [data], left=left, right=right, value=location, at=at)


def step_process(data, left, right, location, at):
    if not _is_values(data):
        return data
    return run_single(lambda ops: _lower_step(ops + [left, right, location, at]), [data])


step_process._gm_lower = _lower_step


def _lower_classify(args):
    return Node("classify", args[:1], bins=args[1], right=args[2])


def classify_process(data, bins, right):
    if not _is_values(data):
        return data
    return run_single(lambda ops: _lower_classify(ops + [bins, right]), [data])


classify_process._gm_lower = _lower_classify


def _lower_reclassify(args):
    kwargs = args[1]
    return Node("reclassify", args[:1], dtype=kwargs["dtype"], fillvalue=kwargs["fillvalue"],
                data=kwargs["data"], select=kwargs["select"])


def reclassify_process(store_data, process_kwargs):
    if not _is_values(store_data):
        return store_data
    return run_single(lambda ops: _lower_reclassify(ops + [process_kwargs]), [store_data])


reclassify_process._gm_lower = _lower_reclassify
