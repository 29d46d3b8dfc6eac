"""ctypes wrapper of ``polyfill.c`` (GDAL scanline fill restatement) -- ORACLE,
test infrastructure only; see the C file for the algorithm and its provenance."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "polyfill.c")
_LIB = os.path.join(_HERE, "_build", "libpolyfill.so")
_lib = None


def build():
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _LIB, _SRC, "-lm"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.gm_oracle_spans.restype = ctypes.c_int64
    return _lib


def soup(polygons):
    """polygons: list of lists of rings (each ring an (n, 2) array-like, closed or
    not) -> (xy, ring_offsets, poly_offsets) in the CSR layout of GmPolygons."""
    xy, ring_offsets, poly_offsets = [], [0], [0]
    n = 0
    for rings in polygons:
        for ring in rings:
            ring = np.asarray(ring, dtype=np.float64).reshape(-1, 2)
            if len(ring) and not np.array_equal(ring[0], ring[-1]):
                ring = np.vstack([ring, ring[:1]])
            xy.append(ring)
            n += len(ring)
            ring_offsets.append(n)
        poly_offsets.append(len(ring_offsets) - 1)
    xy = np.ascontiguousarray(np.concatenate(xy) if xy else np.zeros((0, 2)), dtype=np.float64)
    return xy, np.asarray(ring_offsets, np.int64), np.asarray(poly_offsets, np.int64)


def geotransform(bbox, height, width):
    x1, y1, x2, y2 = bbox
    return np.array([x1, (x2 - x1) / width, 0.0, y2, 0.0, (y1 - y2) / height], dtype=np.float64)


def burn_index(polygons, bbox, height, width, unlabelled=np.iinfo(np.int32).max):
    """(height, width) int32: index of the LAST polygon whose interior covers the
    cell centre, ``unlabelled`` elsewhere."""
    xy, ro, po = soup(polygons)
    gt = geotransform(bbox, height, width)
    labels = np.full((height, width), unlabelled, dtype=np.int32)
    _load().gm_oracle_burn_index(
        xy.ctypes.data_as(ctypes.c_void_p), ro.ctypes.data_as(ctypes.c_void_p),
        po.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(po) - 1),
        gt.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(height), ctypes.c_int(width),
        labels.ctypes.data_as(ctypes.c_void_p))
    return labels


def spans(polygons, bbox, height, width):
    """(n, 4) int64 rows (polygon, row, x0, x1 inclusive)."""
    xy, ro, po = soup(polygons)
    gt = geotransform(bbox, height, width)
    args = (xy.ctypes.data_as(ctypes.c_void_p), ro.ctypes.data_as(ctypes.c_void_p),
            po.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(po) - 1),
            gt.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(height), ctypes.c_int(width))
    n = _load().gm_oracle_spans(*args, ctypes.c_void_p(), ctypes.c_int64(0))
    out = np.zeros((max(n, 1), 4), dtype=np.int64)
    _load().gm_oracle_spans(*args, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n))
    return out[:n]


def rasterize(polygons, bbox, height, width, values=None):
    """Oracle of utils.rasterize_geoseries for polygon input (utils.py:638-756)."""
    idx = burn_index(polygons, bbox, height, width, unlabelled=-1)
    if values is None:
        return (idx >= 0)[np.newaxis], None
    values = np.asarray(values)
    if values.dtype.kind == "f":
        dtype, nodata = np.float64, np.finfo(np.float64).max
    else:
        dtype, nodata = np.int32, np.iinfo(np.int32).max
    out = np.full((1, height, width), nodata, dtype=dtype)
    out[0][idx >= 0] = values.astype(dtype)[idx[idx >= 0]]
    return out, nodata
