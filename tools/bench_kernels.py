"""Per-kernel throughput of every hot-path row on device-resident inputs.

    python tools/bench_kernels.py [--only smooth,zonal,...] [--scale 1.0] [--iters 10]

Each line: op, ms per call, Gpixel/s, algorithmic GB/s and the fraction of
MEASURED_PEAKS.json hbm_gbs.  Inputs are torch tensors wrapped as DeviceArrays;
timing uses CUDA events on the stream the kernels are launched on.  Sizes are
scaled-down versions of BASELINE.json configs 3-5 (all >> the 126 MB L2).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default="")
    ap.add_argument("--temporal-frames", type=int, default=64, help="365 = cfg5")
    ap.add_argument("--temporal-size", type=int, default=4096, help="8192 = cfg5 (98 GB of float32 at 365 frames)")
    ap.add_argument("--temporal-stats", default="sum,max,mean,median")
    ap.add_argument("--temporal-dtype", default="f4", choices=["f4", "i2"])
    ap.add_argument("--zonal-size", type=int, default=0, help="raster edge of the zonal leg (40000 = cfg4)")
    ap.add_argument("--zonal-grid", type=int, default=0, help="polygons per side (316 = cfg4)")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)

    import torch

    from dask_geomodeling_b200 import _native, geometry, raster, utils, workloads
    from dask_geomodeling_b200.core import fusion

    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    results = []

    def wrap(t):
        dtype = {torch.float32: "f4", torch.float64: "f8", torch.int16: "i2", torch.int32: "i4",
                 torch.uint8: "u1", torch.int64: "i8", torch.bool: "?"}[t.dtype]
        return _native.DeviceArray(tuple(t.shape), dtype, ptr=t.data_ptr(), owner=t)

    def measure(name, fn, pixels, nbytes, iters=None):
        if only and not any(name.startswith(o) for o in only):
            return
        iters = iters or args.iters
        with _native.use_stream(stream.cuda_stream), fusion.device_resident():
            for _ in range(3):
                out = fn()
            torch.cuda.synchronize()
            before = _native.launch_count()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(stream)
            for _ in range(iters):
                out = fn()
            stop.record(stream)
            torch.cuda.synchronize()
            launches = (_native.launch_count() - before) / iters
            del out
        ms = start.elapsed_time(stop) / iters
        gbs = nbytes / ms / 1e6
        row = {"op": name, "ms": round(ms, 4), "gpx_s": round(pixels / ms / 1e6, 2),
               "alg_gb_s": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3),
               "launches": launches}
        results.append(row)
        print(json.dumps(row), flush=True)

    def dem(h, w, nodata_fraction=0.01):
        y = torch.arange(h, device="cuda", dtype=torch.float32)[:, None]
        x = torch.arange(w, device="cuda", dtype=torch.float32)[None, :]
        z = 50 * torch.sin(x / 17.0) + 30 * torch.cos(y / 11.0) + 0.05 * x + torch.randn(h, w, device="cuda") + 100
        z[torch.rand(h, w, device="cuda") < nodata_fraction] = workloads.F32_MAX
        return z[None].contiguous()

    n = int(16384 * args.scale)
    nodata = workloads.F32_MAX

    # ---- stencils (cfg3 shapes at 16k x 16k) -------------------------------------------
    if not only or any(o.startswith(("smooth", "movingmax", "hillshade", "dilate")) for o in only):
        z = dem(n, n)
        zd = wrap(z)
        px = (n - 10) * (n - 10)
        measure("smooth_size5_f32", lambda: raster.Smooth.process(
            {"values": zd, "no_data_value": nodata}, dict(smooth_mode="exact", fill=0, size=[5.0, 5.0])),
            px, px * 8)
        measure("movingmax_11_f32", lambda: raster.MovingMax.process(
            {"values": zd, "no_data_value": nodata}, 11), px, px * 8)
        measure("movingmax_5_f32", lambda: raster.MovingMax.process(
            {"values": zd, "no_data_value": nodata}, 5), px, px * 8)
        measure("hillshade_f32", lambda: raster.HillShade.process(
            {"values": zd, "no_data_value": nodata},
            dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0)), n * n, n * n * 5)
        cls = torch.randint(0, 6, (1, n, n), device="cuda", dtype=torch.uint8)
        cd = wrap(cls)
        measure("dilate_u8_3values", lambda: raster.Dilate.process(
            {"values": cd, "no_data_value": 255}, [3, 1, 5]), n * n, n * n * 2)
        del z, zd, cls, cd
        torch.cuda.empty_cache()

    # ---- temporal (cfg5 shape scaled: 64 x 4096 x 4096) ------------------------------------
    if not only or any(o.startswith(("temporal", "cumulative")) for o in only):
        T, m = args.temporal_frames, int(args.temporal_size * args.scale)
        tdtype = args.temporal_dtype
        tnodata = nodata if tdtype == "f4" else 32767
        stack = torch.empty(T, m, m, device="cuda", dtype=torch.float32 if tdtype == "f4" else torch.int16)
        for t in range(T):   # frame by frame: no second stack-sized temporary
            if tdtype == "f4":
                stack[t].uniform_(0, 100)
            else:
                stack[t].random_(0, 3000)
            stack[t][torch.rand(m, m, device="cuda") < 0.03] = tnodata
        sd = wrap(stack)
        from datetime import datetime, timedelta

        times = [datetime(2000, 1, 1) + timedelta(days=i) for i in range(T)]
        for stat in [x for x in args.temporal_stats.split(",") if x]:
            # dtype_for_statistic (utils.py:826-845): min/max keep the dtype, sums of integers are int32
            out_dtype = "f4" if tdtype == "f4" or stat in ("mean", "median") else ("i2" if stat in ("min", "max") else "i4")
            kwargs = dict(mode="vals", start=times[-1], stop=None, frequency=None, timezone=None,
                          closed=None, label=None, dtype=out_dtype, statistic=stat)
            item = 4 if tdtype == "f4" else 2
            measure("temporal_%s_%s_T%d" % (stat, "f32" if tdtype == "f4" else "i16", T),
                    lambda kwargs=kwargs: raster.TemporalAggregate.process(
                        kwargs, {"time": times}, {"values": sd, "no_data_value": tnodata}),
                    T * m * m, T * m * m * item + m * m * np.dtype(out_dtype).itemsize, iters=max(2, args.iters // 2))
        del stack, sd
        torch.cuda.empty_cache()

    # ---- zonal statistics (cfg4 scaled: 16k x 16k, 128 x 128 polygons) ---------------------------
    if not only or any(o.startswith(("zonal", "rasterize")) for o in only):
        if args.zonal_size:
            n = args.zonal_size
        r = torch.empty(1, n, n, device="cuda")
        rows = max(1, (1 << 27) // n)
        for a in range(0, n, rows):   # chunked: bounded temporaries next to a 6.4 GB raster
            block = r[0, a:a + rows]
            block.uniform_(0, 100)
            block[torch.rand(block.shape, device="cuda") < 0.02] = nodata
        rd = wrap(r)
        g = args.zonal_grid or int(128 * args.scale)
        polys = workloads.cfg4_polygons(n, g)   # the generator bench.py and the parity tests use
        bbox = (0, 0, n, n)
        soup = utils.PolygonSoup(polys).to_device()   # CSR build (6 us / polygon) + upload kept out of the timing
        for stat, q in (("mean", None), ("max", None), ("percentile", 90.0), ("median", None)):
            measure("zonal_%s_f32_%dpolys" % (stat if q is None else "p90", len(polys)),
                    lambda stat=stat, q=q: geometry.aggregate.aggregate_polygons(
                        soup, rd, nodata, bbox, workloads.PROJECTION, None, stat, q),
                    n * n, n * n * 4, iters=max(2, args.iters // 3))
        import pandas as pd

        ids = pd.Series(np.arange(len(polys)))
        series = pd.Series(polys, dtype=object)
        measure("rasterize_int32_%dpolys" % len(polys),
                lambda: utils.rasterize_geoseries(series, bbox, workloads.PROJECTION, n, n, values=ids, soup=soup),
                n * n, n * n * 4, iters=max(2, args.iters // 3))
        del r, rd
        torch.cuda.empty_cache()

    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
