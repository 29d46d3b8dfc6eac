"""Oracle evaluation of the BASELINE.json chains (block by block, NumPy).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Each function applies
the oracle restatement of every block of the corresponding view in
``dask_geomodeling_b200/workloads.py``, in the order the reference's graph
executes them.
"""
import numpy as np

from . import raster as R

F32_MAX = float(np.finfo(np.float32).max)


def cfg1(a, b):
    """Add -> Multiply -> Greater -> Clip -> Mask."""
    s = R.elementwise("add", "float32", F32_MAX, (a, F32_MAX), (b, F32_MAX))
    p = R.elementwise("multiply", "float32", F32_MAX, s, 0.5)
    g = R.elementwise("greater", "bool", None, p, 40.0)
    c = R.clip(p[0], p[1], g[0], g[1])
    return R.mask(c[0], c[1], 1)


def cfg2(ints, floats, pairs):
    """Reclassify -> Clip -> Step -> IsData; returns (isdata, step) pairs."""
    fill = np.iinfo(np.int64).max
    r = R.reclassify(ints, 32767, pairs, True, "int64", fill)
    c = R.clip(floats, F32_MAX, r[0], r[1])
    st = R.step(c[0], c[1], 0, 1, 50.0, 0.5)
    return R.is_data(st[0], st[1]), st
