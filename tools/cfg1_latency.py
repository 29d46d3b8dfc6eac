"""BASELINE.json configs[0]: the 1024 x 1024 float32 Add/Multiply/Greater/Clip/Mask chain through
``view.get_data`` on host arrays (request latency), next to the CPU oracle port."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dask_geomodeling_b200 import _native, workloads
    from oracle import workloads as oracle_workloads

    size = 1024
    a, b = workloads.cfg1_arrays(size)
    view = workloads.cfg1_view(a, b)
    request = workloads.request(size, size)
    for _ in range(5):
        got = view.get_data(**request)
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        got = view.get_data(**request)
    gpu = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(5):
        expected, _ = oracle_workloads.cfg1(a, b)
    cpu = (time.perf_counter() - t0) / 5
    assert np.array_equal(got["values"], expected)
    print("cfg1 1024x1024 get_data: %.3f ms per request (%.2f Gpx/s); CPU port %.1f ms (%.0fx)" % (
        gpu * 1e3, size * size / gpu / 1e9, cpu * 1e3, cpu / gpu))
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        view.get_data(**request)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
