"""Where the time of a striped zonal call goes (torchrun, N ranks): CUDA events around the
stripe kernels, the collectives and the finalisation of parallel.zonal_striped."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["GM_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from dask_geomodeling_b200 import _native, parallel, utils, workloads

    n, g = 40000, 316
    nodata = workloads.F32_MAX
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r0, r1 = parallel.stripe_rows(n, world)[rank]
    stripe = torch.rand(1, r1 - r0, n, device="cuda") * 100
    soup = utils.PolygonSoup(workloads.cfg4_polygons(n, g)).to_device()
    bbox = (0, 0, n, n)
    marks = []
    parallel.TRACE = lambda name: marks.append((name, _mark(torch, stream)))
    with _native.use_stream(stream.cuda_stream):
        for stat, q in (("mean", None), ("percentile", 90.0)):
            for it in range(6):
                marks.clear()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                parallel.TRACE("start")
                parallel.zonal_striped(soup, stripe, nodata, bbox, n, (r0, r1), stat, q)
                parallel.TRACE("end")
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3
            if rank == 0:
                steps = [(b[0], round(a[1].elapsed_time(b[1]), 3)) for a, b in zip(marks[:-1], marks[1:])]
                print(json.dumps({"stat": stat, "n_gpus": world, "wall_ms": round(wall, 3), "gpu_ms_between_marks": steps}))
    if world > 1:
        dist.destroy_process_group()


def _mark(torch, stream):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


if __name__ == "__main__":
    main()
