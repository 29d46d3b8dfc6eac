"""Synthetic inputs and views of the BASELINE.json configurations.

Shared by ``bench.py`` and the parity tests so that the measured workload and
the tested workload are the same code.  Sizes are parameters: the tests run
the same views at sizes the CPU oracle finishes in seconds.
Input definitions follow SURVEY.md section 8(d).
"""
import numpy as np

PROJECTION = "EPSG:28992"
F32_MAX = float(np.finfo(np.float32).max)


def _with_nodata(rng, array, nodata, fraction):
    mask = rng.random(array.shape) < fraction
    array[mask] = nodata
    return array


def cfg1_arrays(size=1024, seed=42):
    """Two float32 rasters, uniform(0, 100), 5 % no data (= float32 max)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 100, (1, size, size)).astype(np.float32)
    b = rng.uniform(0, 100, (1, size, size)).astype(np.float32)
    return _with_nodata(rng, a, F32_MAX, 0.05), _with_nodata(rng, b, F32_MAX, 0.05)


def cfg2_arrays(size=16384, seed=43, chunk=2048):
    """int16 classes 0..49 with 5 % no data (32767) and a float32 raster."""
    rng = np.random.default_rng(seed)
    ints = np.empty((1, size, size), dtype=np.int16)
    floats = np.empty((1, size, size), dtype=np.float32)
    for r0 in range(0, size, chunk):  # chunked: bounded temporaries at 16k x 16k
        r1 = min(r0 + chunk, size)
        shape = (1, r1 - r0, size)
        block = rng.integers(0, 50, shape, dtype=np.int16)
        ints[:, r0:r1] = _with_nodata(rng, block, 32767, 0.05)
        fblock = rng.uniform(0, 100, shape).astype(np.float32)
        floats[:, r0:r1] = _with_nodata(rng, fblock, F32_MAX, 0.05)
    return ints, floats


def source(array, nodata, **kwargs):
    from .raster import MemorySource

    h = array.shape[-2]
    return MemorySource(array, nodata, PROJECTION, pixel_size=1.0, pixel_origin=(0, h), **kwargs)


def request(height, width, **extra):
    req = dict(mode="vals", bbox=(0, 0, width, height), width=width, height=height,
               projection=PROJECTION)
    req.update(extra)
    return req


def cfg1_view(a, b):
    """Add -> Multiply -> Greater -> Clip -> Mask (uint8 out); 4+4 in, 1 out = 9 B/px."""
    from .raster import Add, Clip, Greater, Mask, Multiply

    s = Add(source(a, F32_MAX), source(b, F32_MAX))
    p = Multiply(s, 0.5)
    g = Greater(p, 40.0)
    return Mask(Clip(p, g), 1)


CFG2_PAIRS = [[k, 10 * k] for k in range(0, 50, 2)]


def cfg2_views(ints, floats):
    """Reclassify(int16, select) -> Clip(f32 by it) -> Step -> IsData.

    Returns (IsData view [bool, 2+4 in + 1 out = 7 B/px], Step view [float32,
    2+4 in + 4 out = 10 B/px])."""
    from .raster import Clip, IsData, Reclassify, Step

    r = Reclassify(source(ints, 32767), CFG2_PAIRS, select=True)
    st = Step(Clip(source(floats, F32_MAX), r), left=0, right=1, value=50.0, at=0.5)
    return IsData(st), st


def cfg4_rings(size, grid, seed=7):
    """BASELINE.json configs[3] polygons as SURVEY.md section 8(d) defines them: one convex-ish
    6-12-gon per cell of a ``grid`` x ``grid`` partition of a ``size`` x ``size`` raster.  The
    vertices sit at jittered, roughly even angles on the cell's boundary scaled by 0.93-0.995
    (95 % of the polygons: inside their cell) or 1.04-1.12 (5 %: overlapping their neighbours),
    so that the mean area is ~0.8 of a cell (12.7 k px at 40000 / 316) and the polygons cover
    >= 75 % of the raster; coordinates are rounded to 3 decimals + 0.0137 (never on k + 0.5).
    Returns the list of exterior rings, (k, 2) float64 arrays, in row-major cell order."""
    rng = np.random.default_rng(seed)
    cell = size / grid
    n = grid * grid
    counts = rng.integers(6, 13, n)
    overlapping = rng.random(n) < 0.05
    scale = np.where(overlapping, rng.uniform(1.04, 1.12, n), rng.uniform(0.93, 0.995, n))
    phase = rng.uniform(0, 2 * np.pi, n)
    jitter = rng.uniform(0.15, 0.85, (n, 12))
    shrink = rng.uniform(0.97, 1.0, (n, 12))
    ii, jj = np.divmod(np.arange(n), grid)
    cx, cy = (jj + 0.5) * cell, (ii + 0.5) * cell
    rings = [None] * n
    for k in range(6, 13):
        sel = np.nonzero(counts == k)[0]
        ang = (np.arange(k)[None, :] + jitter[sel, :k]) * (2 * np.pi / k) + phase[sel, None]
        cos, sin = np.cos(ang), np.sin(ang)
        reach = 0.5 * cell / np.maximum(np.abs(cos), np.abs(sin))   # centre -> cell boundary
        rad = reach * scale[sel, None] * shrink[sel, :k]
        xy = np.stack([cx[sel, None] + rad * cos, cy[sel, None] + rad * sin], axis=2)
        xy = np.round(xy, 3) + 0.0137
        for row, index in enumerate(sel):
            rings[index] = xy[row]
    return rings


def ring_areas(rings):
    """Shoelace areas of exterior rings (same units as the coordinates, squared)."""
    out = np.empty(len(rings))
    for i, r in enumerate(rings):
        x, y = r[:, 0], r[:, 1]
        out[i] = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return out


def cfg4_polygons(size, grid, seed=7):
    from . import utils

    return [utils.Polygon(ring) for ring in cfg4_rings(size, grid, seed)]
