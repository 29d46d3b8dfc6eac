import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from dask_geomodeling_b200 import _native, utils, workloads
from dask_geomodeling_b200.geometry import aggregate
n, g = 40000, 316
soup = utils.PolygonSoup(workloads.cfg4_polygons(n, g)).to_device()
for tdt, name, nodata in ((torch.int16, "i2", 32767), (torch.uint8, "u1", 255)):
    r = torch.randint(0, 100, (1, n, n), device="cuda", dtype=tdt)
    r[torch.rand(1, n, n, device="cuda") < 0.02] = nodata
    rd = _native.DeviceArray((1, n, n), name, ptr=r.data_ptr(), owner=r)
    torch.cuda.synchronize()
    for stat, q in (("mean", None), ("max", None), ("percentile", 90.0)):
        f = lambda: aggregate.aggregate_polygons(soup, rd, nodata, (0, 0, n, n), workloads.PROJECTION, None, stat, q)
        for _ in range(2): f()
        t0 = time.perf_counter()
        for _ in range(5): f()
        dt = (time.perf_counter() - t0) / 5
        print("%s %-10s %.2f ms = %.0f Gpx/s" % (name, stat, dt * 1e3, n * n / dt / 1e9), flush=True)
    del r, rd
    torch.cuda.empty_cache()
