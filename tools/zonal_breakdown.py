"""Where a gm_zonal_stats call spends its time at cfg4 size: the whole Python call, the C call
alone, and (with ncu) the kernels.  Usage: python tools/zonal_breakdown.py [--scale 2.4414]"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=2.4414)
    args = ap.parse_args()
    import torch

    from dask_geomodeling_b200 import _native, geometry, utils, workloads
    from dask_geomodeling_b200.geometry import aggregate

    n = int(16384 * args.scale)
    g = int(128 * args.scale)
    r = torch.rand(1, n, n, device="cuda") * 100
    r[torch.rand(1, n, n, device="cuda") < 0.02] = workloads.F32_MAX
    rd = _native.DeviceArray((1, n, n), "f4", ptr=r.data_ptr(), owner=r)
    soup = utils.PolygonSoup(workloads.cfg4_polygons(n, g)).to_device()
    bbox = (0, 0, n, n)
    for stat, q in (("mean", None), ("max", None), ("percentile", 90.0)):
        calls = []
        orig = _native.lib().gm_zonal_stats

        def run():
            return aggregate.aggregate_polygons(soup, rd, workloads.F32_MAX, bbox, workloads.PROJECTION, None, stat, q)

        for _ in range(3):
            run()
        _native.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            run()
        _native.synchronize()
        whole = (time.perf_counter() - t0) / 10
        # the C call alone: same arguments, no Python bookkeeping around it
        import ctypes as C
        from dask_geomodeling_b200.raster._program import sentinel
        polys = soup.as_struct()
        geo = (C.c_double * 6)(*utils.GeoTransform.from_bbox(bbox, n, n))
        holder, nodata_ptr = _native.scalar_ptr(sentinel(np.dtype("f4"), workloads.F32_MAX), np.dtype("f4"))
        desc = aggregate._frame_descriptor(rd, 0)
        out = _native.pinned_empty((soup.n_polygons,), np.float32)
        covered = _native.pinned_empty((soup.n_polygons,), np.int64)
        lib = _native.lib()
        def c_call():
            _native.check(lib.gm_zonal_stats(C.byref(desc), nodata_ptr, 1, C.byref(polys), geo,
                                             aggregate._STAT_CODES[stat], float(q or 0.0), None, 0, n,
                                             out.ctypes.data, covered.ctypes.data, None, _native.current_stream()))
        for _ in range(3):
            c_call()
        t0 = time.perf_counter()
        for _ in range(10):
            c_call()
        c_only = (time.perf_counter() - t0) / 10
        print("%-10s whole call %.3f ms = %.0f Gpx/s   C call alone %.3f ms" % (
            stat, whole * 1e3, n * n / whole / 1e9, c_only * 1e3), flush=True)


if __name__ == "__main__":
    main()
