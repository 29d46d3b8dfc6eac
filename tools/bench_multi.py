"""Strong-scaling benchmark of the sharded stencil and zonal paths (BASELINE.json configs 3
and 4) over the GPUs of one box.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_multi.py \
        [--dem 32768] [--raster 40000] [--polygons 100000] [--iters 5]

The whole raster is cut in row stripes, one per rank, generated directly in HBM.  A step is
halo exchange (ncclSend/Recv) + kernel for the stencils, and stripe reduce + all-reduce
(+ segment routing for p90) for the zonal statistics.  Times are CUDA-event times on the
launching stream, max over ranks; rank 0 prints one JSON line per operation.
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
    ap.add_argument("--dem", type=int, default=32768)
    ap.add_argument("--raster", type=int, default=40000)
    ap.add_argument("--polygons", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--profile-p90", action="store_true", help="cProfile of the striped p90 call on rank 0")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["GM_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from dask_geomodeling_b200 import _native, parallel, raster, utils, workloads

    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    nodata = workloads.F32_MAX
    stream = torch.cuda.Stream()   # a real (non-default) stream: kernels, NCCL and events share it
    torch.cuda.set_stream(stream)

    def measure(name, fn, pixels, nbytes):
        if only and not any(name.startswith(o) for o in only):
            return
        with _native.use_stream(stream.cuda_stream):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            start.record(stream)
            for _ in range(args.iters):
                fn()
            stop.record(stream)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / args.iters * 1e3
        ms = torch.tensor([start.elapsed_time(stop) / args.iters, wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            t = float(ms[0])
            print(json.dumps({
                "op": name, "n_gpus": world, "ms": round(t, 3), "wall_ms": round(float(ms[1]), 3),
                "gpx_s": round(pixels / t / 1e6, 1), "alg_gb_s": round(nbytes / t / 1e6, 1),
                "frac_of_measured_hbm_x_gpus": round(nbytes / t / 1e6 / (peak * world), 3)}), flush=True)

    # ---- cfg3: Smooth / MovingMax / HillShade on a DEM sharded in row stripes ---------------
    if not only or any(o.startswith(("smooth", "movingmax", "hillshade")) for o in only):
        n = args.dem
        r0, r1 = parallel.stripe_rows(n, world)[rank]
        gen = torch.Generator(device="cuda").manual_seed(100 + rank)
        y = torch.arange(r0, r1, device="cuda", dtype=torch.float32)[:, None]
        x = torch.arange(n, device="cuda", dtype=torch.float32)[None, :]
        dem = 50 * torch.sin(x / 17.0) + 30 * torch.cos(y / 11.0) + 0.05 * x + 100
        dem += torch.randn(r1 - r0, n, device="cuda", generator=gen)
        dem[torch.rand(r1 - r0, n, device="cuda", generator=gen) < 0.01] = nodata
        dem = dem[None].contiguous()
        px = n * n
        lw = parallel.smooth_halo(5.0)

        def stored(halo_rows, halo_cols):  # the stripe kept in HBM together with its halo
            return parallel.pad_columns(parallel.exchange_halo(dem, halo_rows, nodata), halo_cols, nodata)

        h = stored(lw, 5)
        measure("smooth_size5", lambda: parallel.stencil_haloed(
            raster.Smooth.process, h, nodata, lw, 5,
            dict(smooth_mode="exact", fill=0, size=[5.0, 5.0], margin=(lw, 5))), px, px * 8)
        h = stored(5, 5)
        measure("movingmax_11", lambda: parallel.stencil_haloed(
            raster.MovingMax.process, h, nodata, 5, 5, 11), px, px * 8)
        h = stored(1, 1)
        measure("hillshade", lambda: parallel.stencil_haloed(
            raster.HillShade.process, h, nodata, 1, 1,
            dict(resolution=(1.0, 1.0), altitude=45.0, azimuth=315.0, fill=0)), px, px * 5)
        del dem, h
        torch.cuda.empty_cache()

    # ---- cfg4: zonal mean / max / p90 of polygons over a raster sharded in row stripes ------
    if not only or any(o.startswith("zonal") for o in only):
        n = args.raster
        r0, r1 = parallel.stripe_rows(n, world)[rank]
        gen = torch.Generator(device="cuda").manual_seed(200 + rank)
        r = torch.rand(1, r1 - r0, n, device="cuda", generator=gen) * 100
        r[torch.rand(1, r1 - r0, n, device="cuda", generator=gen) < 0.02] = nodata
        g = int(round(args.polygons ** 0.5))
        cell = n / g
        rng = np.random.default_rng(7)  # same polygons on every rank
        k = rng.integers(6, 13, g * g)
        polys = []
        for idx in range(g * g):
            i, j = divmod(idx, g)
            cx, cy = (j + 0.5) * cell, (i + 0.5) * cell
            ang = np.sort(rng.uniform(0, 2 * np.pi, k[idx]))
            rad = cell * rng.uniform(0.40, 0.55, k[idx])
            ring = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
            polys.append(utils.Polygon(np.round(ring, 3) + 0.0137))
        soup = utils.PolygonSoup(polys).to_device()   # polygons resident in HBM, as a server would keep them
        bbox = (0, 0, n, n)
        px = n * n
        for stat, q in (("mean", None), ("max", None), ("percentile", 90.0)):
            measure("zonal_%s_%dpolys" % ("p90" if q else stat, len(polys)),
                    lambda stat=stat, q=q: parallel.zonal_striped(soup, r, nodata, bbox, n, (r0, r1), stat, q),
                    px, px * 4)
        if args.profile_p90:
            import cProfile
            import pstats

            pr = cProfile.Profile()
            pr.enable()
            for _ in range(5):
                parallel.zonal_striped(soup, r, nodata, bbox, n, (r0, r1), "percentile", 90.0)
            pr.disable()
            if rank == 0:
                pstats.Stats(pr).sort_stats("tottime").print_stats(22)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
