"""CPU arm of ``bench.py``: BASELINE.json configs[1] and configs[3] on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Nothing under
``dask_geomodeling_b200/`` imports this module.

Two ways to run the chain Reclassify -> Clip -> Step -> IsData on the CPU:

* ``kind = "stub harness"``: the REAL reference ``process`` staticmethods
  (``/root/reference/dask_geomodeling/raster/misc.py:98-123, :309-328, :482-515`` and
  ``raster/elemwise.py:601-607``) imported behind stubbed GDAL/shapely/dask modules by
  ``oracle/refharness.py`` -- available wherever ``GM_REFERENCE_ROOT`` (default /root/reference)
  exists, i.e. in the build container;
* ``kind = "port"``: the oracle's restatement of the same functions (``oracle/raster.py``),
  used where the reference checkout does not exist (the GPU box).

Both are driven the way BASELINE.md prescribes for a box without dask: the 16384 x 16384 request
is cut in tiles of at most 2048 x 2048 cells (what ``RasterTiler(view, 2048)`` does,
reference raster/parallelize.py:43-91) and the tiles are evaluated by a thread pool on all host
cores (NumPy releases the GIL in its inner loops).
"""
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

F32_MAX = float(np.finfo(np.float32).max)


def chain_function(pairs):
    """(function(ints_tile, floats_tile) -> bool tile, kind)."""
    from . import refharness

    if refharness.available():
        ns = refharness.load()
        fill = np.iinfo(np.int64).max
        kwargs = {"dtype": np.dtype("int64").str, "fillvalue": fill, "data": pairs, "select": True}

        def run(ints, floats):
            r = ns.misc.Reclassify.process({"values": ints, "no_data_value": 32767}, kwargs)
            c = ns.misc.Clip.process({"values": floats, "no_data_value": F32_MAX}, r)
            st = ns.misc.Step.process(c, 0, 1, 50.0, 0.5)
            return ns.elemwise.IsData.process(st)["values"]

        return run, "stub harness"

    from . import workloads as ow

    def run(ints, floats):
        (isdata, _), _ = ow.cfg2(ints, floats, pairs)
        return isdata

    return run, "port"


def tiles(height, width, tile=2048):
    return [(r, min(r + tile, height), c, min(c + tile, width))
            for r in range(0, height, tile) for c in range(0, width, tile)]


def time_chain(ints, floats, pairs, steps=1, warmup=0, threads=None, tile=2048):
    """Seconds per step of the tiled chain over the whole (1, H, W) pair; returns
    (list of step times, kind, threads, checksum of the last step)."""
    threads = threads or os.cpu_count() or 1
    run, kind = chain_function(pairs)
    _, height, width = ints.shape
    todo = tiles(height, width, tile)

    def work(t):
        r0, r1, c0, c1 = t
        # tiles are contiguous copies, as the tiles of a RasterTiler request are
        out = run(np.ascontiguousarray(ints[:, r0:r1, c0:c1]), np.ascontiguousarray(floats[:, r0:r1, c0:c1]))
        return int(out.sum())

    times, checksum = [], None
    with ThreadPoolExecutor(threads) as pool:
        for step in range(warmup + steps):
            t0 = time.perf_counter()
            checksum = sum(pool.map(work, todo))
            dt = time.perf_counter() - t0
            if step >= warmup:
                times.append(dt)
    return times, kind, threads, checksum


def time_zonal(size=4096, grid=32, statistic="mean", q=None, threads=None):
    """The reference's ``aggregate_polygons`` (geometry/aggregate.py:113-203) on a cfg4-shaped
    sample: bucketize the polygons into sets of disjoint boxes, burn one label raster per bucket
    (GDAL fill rule restated in oracle/polyfill.c -- GDAL itself is not installed) and reduce with
    scipy.ndimage labelled statistics / measurements.percentile.  The buckets are independent:
    they run on a thread pool (the reference itself runs them one after the other).
    Returns (Gpixel/s, seconds, threads, kind)."""
    from dask_geomodeling_b200 import workloads
    from dask_geomodeling_b200.geometry.aggregate import bucketize
    from . import polyfill
    from . import raster as R

    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(11)
    frame = rng.uniform(0, 100, (size, size)).astype(np.float32)
    frame[rng.random((size, size)) < 0.02] = F32_MAX
    rings = workloads.cfg4_rings(size, grid)
    bbox = (0, 0, size, size)
    polyfill.build()
    t0 = time.perf_counter()
    boxes = [(r[:, 0].min(), r[:, 1].min(), r[:, 0].max(), r[:, 1].max()) for r in rings]

    def bucket(ids):
        labels = polyfill.burn_index([[rings[i]] for i in ids], bbox, size, size)
        labelled = labels != np.iinfo(np.int32).max
        labels[labelled] = np.asarray(ids, dtype=np.int32)[labels[labelled]]
        out, _ = R.zonal_from_labels(frame, F32_MAX, [(labels, ids)], len(rings), statistic, q)
        return ids, out

    result = np.full(len(rings), np.nan, dtype=np.float32)
    with ThreadPoolExecutor(threads) as pool:
        for ids, out in pool.map(bucket, bucketize(boxes)):
            result[ids] = out[ids]
    seconds = time.perf_counter() - t0
    return size * size / seconds / 1e9, seconds, threads, "port"
