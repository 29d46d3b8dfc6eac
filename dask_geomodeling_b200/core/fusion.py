"""Graph optimiser: fuse element-wise sub-graphs into single CUDA launches.

``Block.get_compute_graph`` emits one task per block, exactly like the
reference (core/graphs.py:161-190).  Tasks whose function carries a
``_gm_lower`` attribute (all of raster/elemwise.py and the per-pixel blocks of
raster/misc.py) see the *same* request geometry as their sources
(SURVEY.md Appendix C), so a connected group of them can be evaluated per
pixel in one pass.  ``optimize`` replaces each such group by one task that
keeps the group's root key, so ``compute(graph, name)`` and external caches
keyed on task names keep working.
"""
import contextlib

from .. import _native, _state
from .._compat import config

__all__ = ["optimize", "device_resident", "to_host", "fused_process", "streamed_fused_process"]


def _is_key(arg, graph):
    try:
        return arg in graph
    except TypeError:
        return False


def _fusable(task):
    return type(task) is tuple and len(task) > 0 and hasattr(task[0], "_gm_lower")


def optimize(graph, name):
    """Return a graph in which groups of >= 2 fusable tasks are single tasks."""
    if not config.get("geomodeling.fuse", True):
        return graph
    consumers = {}
    for key, task in graph.items():
        if type(task) is not tuple:
            continue
        for arg in task[1:]:
            if _is_key(arg, graph) and isinstance(arg, str):
                consumers.setdefault(arg, set()).add(key)

    out = dict(graph)
    done = set()

    def visit(root):
        if root in done or root not in graph:
            return
        done.add(root)
        task = graph[root]
        if type(task) is not tuple:
            return
        if not _fusable(task):
            for arg in task[1:]:
                if _is_key(arg, graph) and isinstance(arg, str):
                    visit(arg)
            return
        # grow the group: absorb fusable children all of whose consumers are inside
        group = {root}
        changed = True
        while changed:
            changed = False
            for member in list(group):
                for arg in graph[member][1:]:
                    if (
                        isinstance(arg, str) and _is_key(arg, graph) and arg not in group
                        and _fusable(graph[arg]) and consumers.get(arg, set()) <= group
                    ):
                        group.add(arg)
                        changed = True
        leaves = []

        def describe(key):
            func = graph[key][0]
            spec = []
            for arg in graph[key][1:]:
                if isinstance(arg, str) and _is_key(arg, graph):
                    if arg in group:
                        spec.append(("node", arg))
                    else:
                        if arg not in leaves:
                            leaves.append(arg)
                        spec.append(("leaf", leaves.index(arg)))
                else:
                    spec.append(("lit", arg))
            return func, spec

        if len(group) > 1:
            plan = {"root": root, "nodes": {key: describe(key) for key in group}}
            # `describe` filled `leaves` in first-seen order
            out[root] = (fused_process, plan) + tuple(leaves)
            if root == name:
                sources = [_streamable_source(graph.get(leaf)) for leaf in leaves]
                if sources and all(kw is not None for kw in sources) and _worth_streaming(sources):
                    # the requested key itself, fed by in-memory sources only: upload, evaluate
                    # and download in row chunks on rotating streams (dict literals, not keys)
                    out[root] = (streamed_fused_process, plan) + tuple(sources)
            for key in group:
                if key != root:
                    out.pop(key, None)
                done.add(key)
            for leaf in leaves:
                visit(leaf)
        else:
            for arg in task[1:]:
                if _is_key(arg, graph) and isinstance(arg, str):
                    visit(arg)

    visit(name)
    return out


def build_expression(plan, n_leaves=None, extra_leaves=None):
    """Expression DAG (raster/_program.Node) of a fused group.

    ndarray literals of math / comparison / logic tasks are raster operands
    (raster/elemwise.py:235-299 accepts them like the reference): each becomes an extra
    leaf without a no data value, numbered from ``n_leaves`` on and appended to
    ``extra_leaves``; without that list such a group raises FusionLimit."""
    import numpy as np

    from ..raster._program import FusionLimit, Leaf

    nodes = plan["nodes"]
    built = {}

    def build(key):
        if key in built:
            return built[key]
        func, spec = nodes[key]
        operands = []
        for kind, value in spec:
            if kind == "node":
                operands.append(build(value))
            elif kind == "leaf":
                operands.append(Leaf(value))
            elif isinstance(value, np.ndarray) and getattr(func, "_gm_array_operands", False):
                if extra_leaves is None or n_leaves is None:
                    raise FusionLimit("array operand in a fused group")
                operands.append(Leaf(n_leaves + len(extra_leaves)))
                extra_leaves.append(value)
            else:
                operands.append(value)
        built[key] = func._gm_lower(operands)
        return built[key]

    return build(plan["root"])


STREAM_MIN_PIXELS = 1 << 24      # below this one upload + one launch is as fast
STREAM_CHUNK_PIXELS = 1 << 24    # pixels (all bands) per pipeline chunk


def _streamable_source(task):
    """The process_kwargs of a MemorySource ``vals`` task whose request is pixel-aligned
    with the source (row_step == col_step == 1), else None."""
    from ..raster.sources import RasterSourceBase, window_geometry
    from .. import utils

    if type(task) is not tuple or len(task) != 2 or task[0] is not RasterSourceBase.process:
        return None
    kw = task[1]
    if not isinstance(kw, dict) or kw.get("mode") != "vals" or "array" not in kw:
        return None    # (file sources decode their window first: not part of the chunk pipeline)
    bbox = kw["bbox"]
    if bbox[0] == bbox[2] or bbox[1] == bbox[3] or kw["width"] == 0 or kw["height"] == 0:
        return None
    if not utils.same_projection(kw["projection"], kw["source_projection"]):
        return None
    _, col_step, _, row_step = window_geometry(kw["geo_transform"], bbox, kw["height"], kw["width"])
    if col_step != 1.0 or row_step != 1.0:
        return None
    return kw


def _worth_streaming(sources):
    if not config.get("geomodeling.stream", True):
        return False
    if int(config.get("geomodeling.device-cache-bytes", 0) or 0) > 0:
        return False   # sources stay resident in HBM: nothing to pipeline after the first request
    first = sources[0]
    bands = first["bands"][1] - first["bands"][0]
    if any(kw["bands"][1] - kw["bands"][0] != bands or kw["height"] != first["height"]
           or kw["width"] != first["width"] for kw in sources):
        return False
    return bands * first["height"] * first["width"] >= STREAM_MIN_PIXELS and bands >= 1


def streamed_fused_process(plan, *sources):
    """A fused group that is the requested key and reads only in-memory sources.

    The request is cut in row chunks; chunk c is uploaded (pinned H2D), evaluated (one launch
    of the compiled program) and downloaded (into the pinned result) on stream c % 3, so the
    upload of a chunk overlaps the download of the previous one (PCIe is full duplex) and the
    kernels hide behind both.  Results are identical to the unchunked evaluation: pixels are
    independent and aligned requests crop rows without resampling arithmetic."""
    import numpy as np

    from .. import utils
    from ..raster import _program
    from ..raster.sources import RasterSourceBase, resample_window

    def fallback():
        with _state.device_resident(True):
            payloads = [RasterSourceBase.process(kw) for kw in sources]
            result = fused_process(plan, *payloads)
        return to_host(result)

    first = sources[0]
    b0, b1 = first["bands"]
    bands, height, width = b1 - b0, first["height"], first["width"]
    rows = max(64, STREAM_CHUNK_PIXELS // max(bands * width, 1))
    if height < 2 * rows:
        return fallback()
    leaf_types = [(np.dtype(kw["dtype"]), kw["fillvalue"].item()) for kw in sources]
    try:
        prog, compiler, results = _program.compile_expression([build_expression(plan)], leaf_types)
    except (_program.FusionLimit, KeyError, NotImplementedError, TypeError):
        return fallback()
    out_dtype = results[0].dtype
    out = _native.pinned_empty((bands, height, width), out_dtype)
    if config.get("geomodeling.pin-sources", True):
        for kw in sources:
            _native.pin(kw["array"])
    lib = _native.lib()
    streams = _native.pipeline_streams(3)
    item = np.dtype(out_dtype).itemsize
    for c, r0 in enumerate(range(0, height, rows)):
        r1 = min(r0 + rows, height)
        stream = streams[c % len(streams)]
        with _native.use_stream(stream):
            inputs = [
                resample_window(kw["array"], kw["bands"], utils.GeoTransform(kw["geo_transform"]),
                                kw["fillvalue"].item(), kw["bbox"], height, width, True, row_range=(r0, r1))
                for kw in sources
            ]
            chunk, = _program.run_program(prog, inputs, [out_dtype], (bands, r1 - r0, width), True)
            plane = (r1 - r0) * width * item
            for band in range(bands):
                _native.check(lib.gm_memcpy_d2h_async(
                    out[band, r0:r1].ctypes.data, chunk.ptr + band * plane, plane, stream))
            del inputs, chunk   # freed stream-ordered (cudaFreeAsync on this chunk's stream)
    for stream in streams:
        _native.stream_sync(stream)
    del compiler
    return {"values": out, "no_data_value": results[0].nodata}


# Programs compiled for a fused group, keyed by the group's structure (task names are block
# tokens, so a view asked again -- another tile, another time -- finds its program here;
# lowering an expression costs about as much as launching it on a 1024 x 1024 tile).
_compiled_plans = {}


def _freeze(value):
    import numpy as np

    if isinstance(value, np.ndarray):
        return ("ndarray", value.shape, value.dtype.str, value.tobytes())
    if isinstance(value, dict):
        return tuple(sorted((k, _freeze(v)) for k, v in value.items()))
    if isinstance(value, (list, tuple)):
        return tuple(_freeze(v) for v in value)
    if isinstance(value, float) and value != value:
        return ("nan",)
    hash(value)
    return (type(value).__name__, value)


def _plan_key(plan, leaves):
    """Hashable identity of (expression, leaf dtypes / no data); None when some literal of
    the plan cannot be hashed."""
    try:
        nodes = tuple(
            (name, getattr(func, "__module__", ""), getattr(func, "__qualname__", repr(func)),
             tuple((kind, _freeze(value)) for kind, value in spec))
            for name, (func, spec) in sorted(plan["nodes"].items()))
        types = tuple((str(v.dtype), _freeze(nd)) for v, nd in leaves)
        return (plan["root"], nodes, types)
    except TypeError:
        return None


def _payload_is_raster(data):
    return isinstance(data, dict) and "values" in data


def fused_process(plan, *leaf_data):
    """Evaluate a fused group.  With raster payloads on every leaf the whole
    group is one program; otherwise (None, time or meta payloads) the original
    ``process`` functions are applied one by one, which reproduces the
    reference's pass-through rules without touching pixel data."""
    from ..raster import _program

    nodes = plan["nodes"]
    if all(_payload_is_raster(d) for d in leaf_data):
        try:
            leaves = [(d["values"], d.get("no_data_value")) for d in leaf_data]
            arrays = plan.get("array_operands")
            if arrays is None:
                # ndarray operands of the group (found once per plan; they are literals of
                # the graph, so their identity is part of the plan)
                arrays = []
                expression = build_expression(plan, len(leaves), arrays)
                plan["array_operands"] = arrays
            else:
                expression = None
            leaves += [(a, None) for a in arrays]
            key = _plan_key(plan, leaves)
            compiled = _compiled_plans.get(key) if key is not None else None
            if compiled is None:
                if expression is None:
                    expression = build_expression(plan, len(leaf_data), [])
                compiled = _program.compile_expression(
                    [expression], [(v.dtype, nd) for v, nd in leaves])
                if key is not None:
                    if len(_compiled_plans) >= 256:
                        _compiled_plans.pop(next(iter(_compiled_plans)))
                    _compiled_plans[key] = compiled
            (values, dtype, nodata), = _program.evaluate(None, leaves, _state.keep_on_device(), compiled)
            return {"values": values, "no_data_value": nodata}
        except _program.FusionLimit:
            pass  # too large for one program: evaluate block by block below
        except (KeyError, NotImplementedError, TypeError):
            # an operand form the fused lowering does not know: the blocks' own process
            # functions below decide (and raise the reference's errors where it would)
            pass

    cache = {}

    def run(key):
        if key in cache:
            return cache[key]
        func, spec = nodes[key]
        args = []
        for kind, value in spec:
            if kind == "node":
                args.append(run(value))
            elif kind == "leaf":
                args.append(leaf_data[value])
            else:
                args.append(value)
        cache[key] = func(*args)
        return cache[key]

    try:
        return run(plan["root"])
    finally:
        cache.clear()  # `run` is recursive: break the cycle that would keep the rasters alive


@contextlib.contextmanager
def device_resident():
    """Context in which blocks hand DeviceArrays to each other."""
    with _state.device_resident(bool(config.get("geomodeling.device-resident", True))):
        yield


def to_host(result):
    """Copy device-resident raster payloads of a task result back to numpy."""
    if isinstance(result, dict) and _native.is_device(result.get("values")):
        result = dict(result)
        result["values"] = result["values"].to_host()
    return result
