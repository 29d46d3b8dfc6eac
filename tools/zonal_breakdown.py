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
        print("%-10s whole call %.3f ms = %.0f Gpx/s" % (stat, whole * 1e3, n * n / whole / 1e9), flush=True)


if __name__ == "__main__":
    main()
