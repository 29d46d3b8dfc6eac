"""Where the time of one end-to-end ``get_data`` call on cfg2 goes (host arrays in,
host array out): raw pinned / pageable copy bandwidth, then the stages of the call."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    from dask_geomodeling_b200 import _native, workloads

    lib = _native.lib()

    def t(label, fn, n=3):
        best = 1e9
        for _ in range(n):
            _native.synchronize()
            t0 = time.perf_counter()
            r = fn()
            _native.synchronize()
            best = min(best, time.perf_counter() - t0)
        print("%-44s %8.2f ms" % (label, best * 1e3), flush=True)
        return r

    nbytes = 1 << 30
    dev = _native.DeviceArray((nbytes,), "u1")
    pinned = _native.pinned_empty((nbytes,), "u1")
    pageable = np.ones(nbytes, dtype="u1")
    t("H2D 1 GiB pinned (cudaHostAlloc)", lambda: _native.check(
        lib.gm_memcpy_h2d(dev.ptr, pinned.ctypes.data, nbytes, None)))
    t("H2D 1 GiB pageable", lambda: _native.check(
        lib.gm_memcpy_h2d(dev.ptr, pageable.ctypes.data, nbytes, None)))
    t0 = time.perf_counter()
    ok = _native.pin(pageable)
    print("cudaHostRegister 1 GiB: %s in %.1f ms" % (ok, (time.perf_counter() - t0) * 1e3))
    t("H2D 1 GiB registered", lambda: _native.check(
        lib.gm_memcpy_h2d(dev.ptr, pageable.ctypes.data, nbytes, None)))
    t("D2H 1 GiB pinned", lambda: _native.check(
        lib.gm_memcpy_d2h(pinned.ctypes.data, dev.ptr, nbytes, None)))
    del dev, pinned, pageable

    ints, floats = workloads.cfg2_arrays(size)
    view, _ = workloads.cfg2_views(ints, floats)
    request = workloads.request(size, size)
    t("get_data cold (tokens, pin, compile)", lambda: view.get_data(**request), n=1)
    t("get_data warm", lambda: view.get_data(**request), n=3)
    t("get_compute_graph only", lambda: view.get_compute_graph(**request), n=3)
    import cProfile
    import pstats

    prof = cProfile.Profile()
    prof.enable()
    view.get_data(**request)
    prof.disable()
    pstats.Stats(prof).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
