"""Timeline of ONE striped zonal call on rank 0 (torchrun, N ranks): every kernel / copy with its
start and duration and every CUDA runtime call of the host thread, from CUPTI (torch.profiler), so
that the gaps between the kernels of a 0.5 ms call can be attributed.  Not a timing tool: the
numbers carry the profiler's overhead; `tools/zonal_striped_breakdown.py` has the clean ones."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["GM_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from dask_geomodeling_b200 import _native, parallel, utils, workloads

    n, g = 40000, 316
    fake_world = int(os.environ.get("GM_FAKE_WORLD", world))     # stripe height of a larger job
    nodata = workloads.F32_MAX
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r0, r1 = parallel.stripe_rows(n, fake_world)[rank]
    stripe = torch.rand(1, r1 - r0, n, device="cuda") * 100
    soup = utils.PolygonSoup(workloads.cfg4_polygons(n, g)).to_device()
    bbox = (0, 0, n, n)
    out = {}
    with _native.use_stream(stream.cuda_stream):
        for stat, q in (("mean", None), ("percentile", 90.0)):
            if fake_world != world and stat == "percentile":
                continue      # the boundary exchange needs the real neighbours
            call = lambda: parallel.zonal_striped(soup, stripe, nodata, bbox, n, (r0, r1), stat, q) \
                if fake_world == world else parallel._zonal_striped_device(
                    soup, parallel._as_payload(stripe), nodata, (0, n - r1, n, n - r0), stat, None, None, True)
            for _ in range(4):
                call()
            torch.cuda.synchronize()
            dist.barrier()
            with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                call()
                torch.cuda.synchronize()
            if rank == 0:
                events = [e for e in prof.events() if e.time_range.start > 0]
                t0 = min(e.time_range.start for e in events)
                rows = []
                for e in sorted(events, key=lambda e: e.time_range.start):
                    kind = "gpu" if str(e.device_type).endswith("CUDA") else "cpu"
                    if kind == "cpu" and not (e.name.startswith("cuda") or e.name.startswith("cu") or "nccl" in e.name):
                        continue
                    rows.append([kind, e.name[:70], round(e.time_range.start - t0, 1), round(e.time_range.elapsed_us(), 1)])
                out[stat] = rows
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/zonal_striped_timeline_{}of{}.json".format(world, fake_world), "w") as f:
            json.dump(out, f)
        for stat, rows in out.items():
            print("==", stat)
            for row in rows:
                print("%-4s %-72s %9.1f %8.1f" % tuple(row))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
