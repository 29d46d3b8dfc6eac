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

__all__ = ["optimize", "device_resident", "to_host", "fused_process"]


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


def build_expression(plan):
    """Expression DAG (raster/_program.Node) of a fused group."""
    from ..raster._program import Leaf

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
            else:
                operands.append(value)
        built[key] = func._gm_lower(operands)
        return built[key]

    return build(plan["root"])


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
            root = build_expression(plan)
            leaves = [(d["values"], d.get("no_data_value")) for d in leaf_data]
            (values, dtype, nodata), = _program.evaluate([root], leaves, _state.keep_on_device())
            return {"values": values, "no_data_value": nodata}
        except _program.FusionLimit:
            pass  # too large for one program: evaluate block by block below

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

    return run(plan["root"])


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
