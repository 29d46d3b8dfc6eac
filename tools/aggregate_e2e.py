"""AggregateRaster end to end through the Block API (host raster in a MemorySource, polygons in
a MemoryGeometrySource, ``view.get_data``): where the time of a request goes.
Usage: python tools/aggregate_e2e.py [--size 16384] [--grid 128]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    from dask_geomodeling_b200 import _native, geometry, raster, utils, workloads
    from dask_geomodeling_b200._compat import config

    n, g = args.size, args.grid
    rng = np.random.default_rng(4)
    values = rng.uniform(0, 100, (1, n, n)).astype("f4")
    values[rng.random((1, n, n)) < 0.02] = workloads.F32_MAX
    src = raster.MemorySource(values, workloads.F32_MAX, workloads.PROJECTION, pixel_size=1.0, pixel_origin=(0, n))
    rings = workloads.cfg4_rings(n, g)
    source = geometry.MemoryGeometrySource([utils.Polygon(r) for r in rings], None, workloads.PROJECTION)
    request = dict(mode="intersects", projection=workloads.PROJECTION, geometry=utils.box(0, 0, n, n))
    for stat in ("mean", "p90"):
        view = geometry.AggregateRaster(source=source, raster=src, statistic=stat, max_pixels=4 * n * n)
        with config.set({"geomodeling.raster-limit": 4 * n * n}):
            for _ in range(2):
                result = view.get_data(**request)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                result = view.get_data(**request)
            dt = (time.perf_counter() - t0) / reps
            print("AggregateRaster(%s) %dx%d, %d polygons: %.1f ms per request = %.2f Gpx/s" % (
                stat, n, n, len(rings), dt * 1e3, n * n / dt / 1e9), flush=True)
            if args.profile and stat == "mean":
                pr = cProfile.Profile()
                pr.enable()
                view.get_data(**request)
                pr.disable()
                pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
