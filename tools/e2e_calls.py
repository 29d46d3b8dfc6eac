"""Per-call wall time of view.get_data on cfg2 (host arrays in, host array out), with a
profile of one early (slow) and one late (fast) call."""
import cProfile, os, pstats, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dask_geomodeling_b200 import _native, workloads
from dask_geomodeling_b200._compat import config
size = 16384
ints, floats = workloads.cfg2_arrays(size)
view, _ = workloads.cfg2_views(ints, floats)
request = workloads.request(size, size)
stream = "--single" not in sys.argv
with config.set({"geomodeling.stream": stream}):
    times, allocs = [], []
    for i in range(10):
        prof = cProfile.Profile() if i in (2, 8) else None
        t0 = time.perf_counter()
        if prof: prof.enable()
        r = view.get_data(**request)
        if prof: prof.disable()
        times.append((time.perf_counter() - t0) * 1e3)
        allocs.append(_native.STATS["pinned_allocations"])
        if prof:
            print("---- call", i, "%.1f ms" % times[-1], "gc counts", gc.get_count())
            pstats.Stats(prof).sort_stats("tottime").print_stats(8)
    print(" ".join("%.1f" % t for t in times), "| pinned allocations so far:", allocs, flush=True)
