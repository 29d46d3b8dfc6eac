"""Row-stripe sharding of the raster path over the GPUs of one box.

The reference's only spatial parallelism is ``RasterTiler``
(raster/parallelize.py:43-125): split a ``vals`` request into tiles, compute them
independently, stitch.  This module is that idea across processes -- one process
per GPU, ``torch.distributed`` for the plumbing (NCCL on GPUs, gloo in the CPU
tests) -- with the three exchanges the path needs:

* element-wise / misc / temporal / rasterise: pixels are independent; every rank
  evaluates the request of its own row stripe, no collective (`stripe_request`,
  `get_data_striped`);
* stencils: one neighbour exchange of `halo` rows before the kernel
  (`exchange_halo`, `stencil_striped`); the outer boundary is padded with no
  data exactly like a source does outside its extent;
* zonal statistics: every rank reduces the polygons over its stripe into
  (count, sum, min, max) partials, one all-reduce of N-vectors finishes them
  (`allreduce_partials`, `finalize_partials`, `zonal_striped`); order statistics
  route each polygon's values to its owner rank ``p % world``
  (`exchange_segments`) where `segment_order_statistic` selects them.

Everything that talks to ``torch.distributed`` works on CPU tensors as well, so the
N > 1 logic is covered by world-size-2 gloo tests without a GPU; the compute
calls themselves always go through libgeokernels.so.
"""
import numpy as np

__all__ = [
    "stripe_rows", "stripe_request", "get_data_striped", "get_data_tiled", "exchange_halo", "pad_columns",
    "stencil_striped", "refresh_halo", "stencil_haloed", "smooth_halo", "allreduce_partials", "finalize_partials", "zonal_striped",
    "exchange_segments", "segment_order_statistic",
]


TRACE = None     # tools/zonal_striped_breakdown.py sets a callable(name) to time the phases


def _trace(name):
    if TRACE is not None:
        TRACE(name)


def _dist():
    import torch.distributed as dist

    return dist


def _world(group=None):
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ---------------------------------------------------------------------------
# stripes
# ---------------------------------------------------------------------------


def stripe_rows(height, world):
    """[(r0, r1)] per rank: contiguous row ranges that differ by at most one row."""
    base, extra = divmod(int(height), int(world))
    bounds, r0 = [], 0
    for rank in range(world):
        r1 = r0 + base + (1 if rank < extra else 0)
        bounds.append((r0, r1))
        r0 = r1
    return bounds


def stripe_request(request, rank, world):
    """The request of `rank`'s row stripe and its (r0, r1) in the full request grid.

    Row 0 is the northern edge (utils.GeoTransform.from_bbox), so stripe rows
    [r0, r1) span y in [y2 - r1*dy, y2 - r0*dy]."""
    x1, y1, x2, y2 = request["bbox"]
    height = request["height"]
    r0, r1 = stripe_rows(height, world)[rank]
    dy = (y2 - y1) / height
    sub = dict(request)
    sub["bbox"] = (x1, y2 - r1 * dy, x2, y2 - r0 * dy)
    sub["height"] = r1 - r0
    return sub, (r0, r1)


def get_data_striped(view, group=None, gather=False, **request):
    """Evaluate ``view.get_data(**request)`` as row stripes, one per rank.

    Returns this rank's stripe (a response dict, or None for empty data) and its
    row range; with ``gather=True`` every rank receives the stitched full result
    instead (RasterTiler's stitch, raster/parallelize.py:93-125)."""
    rank, world = _world(group)
    sub, rows = stripe_request(request, rank, world)
    local = None if sub["height"] == 0 else view.get_data(**sub)
    if not gather or world == 1:
        return local, rows
    dist = _dist()
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    filled = [p for p in parts if p is not None]
    if not filled:
        return None, (0, request["height"])
    if "values" not in filled[0]:
        return filled[0], (0, request["height"])
    # a stripe without data (None) becomes 'no data' rows, so that the result keeps its height
    first = filled[0]
    bands, _, width = first["values"].shape
    stitched = []
    for part, (a, b) in zip(parts, stripe_rows(request["height"], world)):
        if part is not None:
            stitched.append(np.asarray(part["values"]))
        elif b > a:
            stitched.append(np.full((bands, b - a, width), first["no_data_value"], dtype=first["values"].dtype))
    values = np.concatenate(stitched, axis=1)
    return {"values": values, "no_data_value": first["no_data_value"]}, (0, request["height"])


def get_data_tiled(view, tile_size, group=None, **request):
    """``RasterTiler`` (raster/parallelize.py) as a multi-GPU scheduler: the request is cut in
    tiles of at most ``tile_size`` cells exactly as the block does, tile i is evaluated by rank
    i mod world (each rank on its own GPU), the tiles are all-gathered and every rank stitches
    the full result.  Suits views whose tiles are independent (element-wise chains, stencils
    whose request margin equals their reach); returns what ``view.get_data(**request)`` returns."""
    from .raster.parallelize import RasterTiler

    rank, world = _world(group)
    tiler = RasterTiler(view, tile_size)
    plan = tiler.get_sources_and_requests(**request)
    kwargs, tiles = plan[0][0], [r for _, r in plan[1:]]
    if kwargs is None:                 # time / meta / point requests are not tiled
        return view.get_data(**request)
    mine = {i: view.get_data(**tiles[i]) for i in range(rank, len(tiles), world)}
    if world > 1:
        gathered = [None] * world
        _dist().all_gather_object(gathered, mine, group=group)
        mine = {}
        for part in gathered:
            mine.update(part)
    return RasterTiler.process(kwargs, *[mine[i] for i in range(len(tiles))])


# ---------------------------------------------------------------------------
# stencils: halo exchange
# ---------------------------------------------------------------------------


def _all_stripes_hold(rows, halo, group):
    """Every rank learns the thinnest stripe: a stripe thinner than the halo cannot serve its
    neighbour, and ALL ranks must refuse together (a rank that raised alone would leave its
    neighbours waiting in the grouped send/recv)."""
    import torch

    dist = _dist()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    thinnest = torch.tensor([int(rows)], dtype=torch.int64, device=dev)
    dist.all_reduce(thinnest, op=dist.ReduceOp.MIN, group=group)
    if int(thinnest[0]) < halo:
        raise ValueError("a stripe of {} rows is thinner than the halo of {} rows".format(int(thinnest[0]), halo))


def exchange_halo(local, halo, fill, group=None):
    """(bands, rows, width) stripe -> (bands, rows + 2*halo, width): `halo` rows of the
    northern neighbour on top, of the southern neighbour below, `fill` at the outer
    boundary.  One grouped send/recv per neighbour (ncclSend/ncclRecv under NCCL)."""
    import torch

    rank, world = _world(group)
    bands, rows, width = local.shape
    out = torch.full((bands, rows + 2 * halo, width), fill, dtype=local.dtype, device=local.device)
    out[:, halo:halo + rows] = local
    if halo == 0 or world == 1:
        return out
    _all_stripes_hold(rows, halo, group)
    dist = _dist()
    # gloo moves host memory: CUDA stripes are staged through the CPU (one-GPU tests);
    # under NCCL the halos go GPU to GPU over NVLink
    staged = local.is_cuda and dist.get_backend(group) != "nccl"
    edge = (lambda t: t.cpu()) if staged else (lambda t: t.contiguous())
    ops, landing = [], []
    if rank > 0:  # northern neighbour
        send = edge(local[:, :halo])
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, recv, rank - 1, group)]
        landing.append((slice(0, halo), recv))
    if rank < world - 1:  # southern neighbour
        send = edge(local[:, rows - halo:])
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, recv, rank + 1, group)]
        landing.append((slice(halo + rows, 2 * halo + rows), recv))
    for work in dist.batch_isend_irecv(ops):
        work.wait()
    for where, recv in landing:
        out[:, where] = recv.to(out.device)
    return out


def refresh_halo(haloed, halo_rows, halo_cols=0, group=None):
    """In-place halo exchange for a stripe that is STORED with its halo.

    `haloed` is (bands, rows + 2*halo_rows, width + 2*halo_cols); its interior is this rank's
    stripe.  The top/bottom `halo_rows` interior rows are sent to the neighbours and their
    edge rows land in the halo rows -- only 2*halo_rows rows move, the stripe itself is not
    copied (what a resident multi-GPU pipeline calls before every stencil step).  Halo rows
    at the outer boundary and the halo columns keep whatever they hold (no data)."""
    rank, world = _world(group)
    if halo_rows == 0 or world == 1:
        return haloed
    rows = haloed.shape[1] - 2 * halo_rows
    if not getattr(haloed, "_stripes_checked", False):     # once per stored stripe, not per step
        _all_stripes_hold(rows, halo_rows, group)
        try:
            haloed._stripes_checked = True
        except AttributeError:
            pass
    works, landing = _post_halo(haloed, halo_rows, group)
    _land_halo(haloed, works, landing)
    return haloed


def _post_halo(haloed, halo_rows, group=None):
    """First half of `refresh_halo`: the sends and receives are posted, nothing is waited for.
    Returns (works, landing) for `_land_halo`."""
    import torch

    rank, world = _world(group)
    dist = _dist()
    rows = haloed.shape[1] - 2 * halo_rows
    staged = haloed.is_cuda and dist.get_backend(group) != "nccl"
    edge = (lambda t: t.cpu()) if staged else (lambda t: t.contiguous())
    ops, landing = [], []
    if rank > 0:
        send = edge(haloed[:, halo_rows:2 * halo_rows])
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, recv, rank - 1, group)]
        landing.append((slice(0, halo_rows), recv))
    if rank < world - 1:
        send = edge(haloed[:, rows:rows + halo_rows])
        recv = torch.empty_like(send)
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, recv, rank + 1, group)]
        landing.append((slice(halo_rows + rows, 2 * halo_rows + rows), recv))
    return (dist.batch_isend_irecv(ops) if ops else []), landing


def _land_halo(haloed, works, landing):
    """Second half: wait for the exchange (the current stream waits under NCCL) and put the
    neighbours' rows into the halo rows."""
    for work in works:
        work.wait()
    for where, recv in landing:
        haloed[:, where] = recv.to(haloed.device)


def pad_columns(values, halo, fill, pitch=1):
    """Add `halo` columns of `fill` on both sides (the x halo a stencil block requests;
    stripes span the full width, so this is always the outer boundary).  With ``pitch`` > 1 the
    row is padded further on the right to a multiple of `pitch` cells (MovingMax stages its tiles
    by TMA when a row is a whole number of 16-byte groups; pass the extra columns as its ``pad``)."""
    import torch

    bands, rows, width = values.shape
    total = width + 2 * halo
    total += (-total) % pitch
    if total == width:
        return values
    out = torch.full((bands, rows, total), fill, dtype=values.dtype, device=values.device)
    out[:, :, halo:halo + width] = values
    return out


def smooth_halo(size_px):
    """Rows a stripe needs from each neighbour for Smooth to equal the whole-raster result:
    the Gaussian radius int(4 * sigma + 0.5) with sigma = size_px / 3 (scipy.ndimage), which
    exceeds the round(size_px) margin the block itself crops (raster/spatial.py:293-295)."""
    return int(4.0 * (float(size_px) / 3.0) + 0.5)


def _as_payload(tensor):
    """torch tensor -> what a block's ``process`` accepts (DeviceArray on CUDA, numpy on CPU)."""
    if tensor.device.type == "cuda":
        from . import _native

        tensor = tensor.contiguous()
        return _native.DeviceArray(tuple(tensor.shape), str(tensor.dtype).replace("torch.", ""),
                                   ptr=tensor.data_ptr(), owner=tensor)
    return tensor.numpy()


def _visible_to_library(tensor):
    """The library launches on its own (non-blocking) stream unless the caller put it on torch's
    (`_native.use_stream`): what torch has queued for `tensor` -- halo rows that just landed, pads --
    must be complete before a kernel of the library reads it."""
    if not getattr(tensor, "is_cuda", False):
        return
    import torch

    from . import _native

    mine = _native.current_stream()
    current = torch.cuda.current_stream(tensor.device)
    if mine is None or int(mine) != int(current.cuda_stream):
        current.synchronize()


def stencil_striped(process, local, no_data_value, halo_rows, halo_cols, *process_args, group=None):
    """Run a stencil block's ``process`` (Smooth, MovingMax, Dilate, HillShade) on a row
    stripe of a raster that is sharded over the ranks.

    `local` is this rank's (bands, rows, width) torch tensor.  The halo the block
    would have requested from its store (raster/spatial.py:27-108) is assembled from the
    neighbours (`exchange_halo`) and from no data at the outer boundary, so the stitched
    result equals the single-GPU result of the whole raster."""
    haloed = exchange_halo(local, halo_rows, no_data_value, group)
    haloed = pad_columns(haloed, halo_cols, no_data_value)
    from .core import fusion

    _visible_to_library(haloed)
    with fusion.device_resident():
        return process({"values": _as_payload(haloed), "no_data_value": no_data_value}, *process_args)


def stencil_haloed(process, haloed, no_data_value, halo_rows, halo_cols, *process_args, group=None,
                   overlap=False):
    """`stencil_striped` for a stripe stored with its halo (see `refresh_halo`): exchange the
    halo rows in place and run the block's ``process`` on the resident array.

    With ``overlap`` (single-band stripes of at least 4 halos) the exchange hides behind the
    kernel: the output rows that do not depend on a neighbour's rows -- all but `halo_rows` at
    either end -- are computed while the 2 * halo_rows rows travel, then the two end strips.  The
    stencil kernels give the same cell the same value whatever window it is computed in (window
    invariance, tests/test_full_size_gpu.py), so the three windows equal the one-call result; on
    the device they are written straight into ONE output raster (`_state.RowWindow`).
    Off by default: measured on 2 B200s (32768 x 32768, profiles/README.md r02s) the overlapped
    form is SLOWER -- Smooth 2.12 against 2.01 ms, MovingMax 1.32 against 1.10 ms, HillShade 0.84
    against 0.82 ms: the NCCL send/recv kernel holds SM resources while it waits for its peer, so
    part of the persistent MovingMax blocks start late, and two more launches follow the interior."""
    from . import _state
    from .core import fusion

    rank, world = _world(group)
    rows = haloed.shape[1] - 2 * halo_rows
    if not (overlap and world > 1 and halo_rows > 0 and haloed.shape[0] == 1 and rows >= 4 * halo_rows
            and hasattr(haloed, "is_cuda")):
        refresh_halo(haloed, halo_rows, halo_cols, group)
        _visible_to_library(haloed)
        with fusion.device_resident():
            return process({"values": _as_payload(haloed), "no_data_value": no_data_value}, *process_args)
    if not getattr(haloed, "_stripes_checked", False):     # once per stored stripe, not per step
        _all_stripes_hold(rows, halo_rows, group)
        try:
            haloed._stripes_checked = True
        except AttributeError:
            pass
    h = halo_rows
    _visible_to_library(haloed)
    works, landing = _post_halo(haloed, h, group)
    target = _state.RowWindow(rows)
    pieces = []

    def window(a, b, r0):       # input rows [a, b) of the stored stripe -> output rows from r0
        target.r0 = r0
        pieces.append(process({"values": _as_payload(haloed[:, a:b]), "no_data_value": no_data_value},
                              *process_args))

    with fusion.device_resident(), _state.into_row_window(target):
        window(h, rows + h, h)                  # interior: needs no neighbour row
        _land_halo(haloed, works, landing)
        _visible_to_library(haloed)
        window(0, 3 * h, 0)                     # the first and the last `h` output rows
        window(rows - h, rows + 2 * h, rows - h)
    result = dict(pieces[0])
    if target.out is not None:
        result["values"] = target.out
    else:                                       # host arrays (gloo on CPU tensors): stitch
        result["values"] = np.concatenate([np.asarray(pieces[1]["values"]), np.asarray(pieces[0]["values"]),
                                           np.asarray(pieces[2]["values"])], axis=1)
    return result


# ---------------------------------------------------------------------------
# zonal statistics: partials + all-reduce
# ---------------------------------------------------------------------------

PARTIAL_DTYPE = np.dtype([("count", "<i8"), ("sum", "<f8"), ("vmin", "<f8"), ("vmax", "<f8")])


def allreduce_partials(partial, covered, group=None, device=None):
    """Combine per-stripe (count, sum, min, max)[N] partials and covered-cell counts
    over all ranks: three all-reduces (sum / min / max) of N-vectors."""
    import torch

    rank, world = _world(group)
    partial = np.asarray(partial, dtype=PARTIAL_DTYPE)
    covered = np.asarray(covered, dtype=np.int64)
    if world == 1:
        return partial.copy(), covered.copy()
    dist = _dist()
    dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    counts = torch.from_numpy(np.stack([partial["count"], covered])).to(dev)
    sums = torch.from_numpy(np.ascontiguousarray(partial["sum"])).to(dev)
    mins = torch.from_numpy(np.ascontiguousarray(partial["vmin"])).to(dev)
    maxs = torch.from_numpy(np.ascontiguousarray(partial["vmax"])).to(dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    out = np.empty(len(partial), dtype=PARTIAL_DTYPE)
    counts = counts.cpu().numpy()
    out["count"], covered = counts[0], counts[1]
    out["sum"], out["vmin"], out["vmax"] = sums.cpu().numpy(), mins.cpu().numpy(), maxs.cpu().numpy()
    return out, covered


def finalize_partials(partial, statistic):
    """float32 statistic per polygon from reduced partials (the arithmetic of
    zonal_finalize_kernel: f64 sum / f64 count -> f32; NaN where nothing was active)."""
    partial = np.asarray(partial, dtype=PARTIAL_DTYPE)
    out = np.full(len(partial), np.nan, dtype=np.float32)
    has = partial["count"] > 0
    if statistic == "count":
        out[has] = partial["count"][has].astype(np.float64).astype(np.float32)
    elif statistic == "sum":
        out[has] = partial["sum"][has].astype(np.float32)
    elif statistic == "mean":
        out[has] = (partial["sum"][has] / partial["count"][has].astype(np.float64)).astype(np.float32)
    elif statistic == "min":
        out[has] = partial["vmin"][has].astype(np.float32)
    elif statistic == "max":
        out[has] = partial["vmax"][has].astype(np.float32)
    else:
        raise ValueError("not a reducible statistic: {}".format(statistic))
    return out


def zonal_striped(geometries, local, no_data_value, bbox, height, rows, statistic, percentile=None,
                  threshold_values=None, group=None):
    """Zonal statistic of a raster that is sharded in row stripes.

    `local` is this rank's (1, r1 - r0, width) stripe (numpy, DeviceArray or CUDA tensor),
    `bbox`/`height` describe the FULL raster grid, `rows` = (r0, r1).  Returns the float32
    statistic per geometry (identical on every rank) and the indices of geometries that
    cover no cell centre anywhere."""
    import ctypes

    from . import _native, utils
    from .geometry.aggregate import _STAT_CODES, _frame_descriptor
    from .raster._program import sentinel

    if hasattr(local, "data_ptr"):
        local = _as_payload(local)
    r0, r1 = rows
    x1, y1, x2, y2 = bbox
    dy = (y2 - y1) / height
    stripe_bbox = (x1, y2 - r1 * dy, x2, y2 - r0 * dy)
    soup = geometries if isinstance(geometries, utils.PolygonSoup) else utils.PolygonSoup(list(geometries))
    n = soup.n_polygons
    if _world(group)[1] == 1:  # one stripe = the whole raster: the single-GPU path, select included
        from .geometry.aggregate import aggregate_polygons

        agg, no_cells = aggregate_polygons(soup, local, no_data_value, stripe_bbox, None, threshold_values,
                                           statistic, percentile)
        return agg[0], no_cells
    order_stat = statistic in ("median", "percentile")
    if not order_stat and n > 0 and _dist().get_backend(group) == "nccl" and _native.is_device(local):
        return _zonal_striped_device(soup, local, no_data_value, stripe_bbox, statistic, threshold_values,
                                     group, r1 > r0)
    if order_stat and n > 0:
        return _zonal_order_striped(soup, local, no_data_value, bbox, height, rows, statistic, percentile,
                                    threshold_values, group)
    partial = np.zeros(n, dtype=PARTIAL_DTYPE)
    partial["vmin"], partial["vmax"] = np.finfo(np.float64).max, -np.finfo(np.float64).max
    covered = np.zeros(n, dtype=np.int64)
    if r1 > r0 and n > 0:
        _, h, w = local.shape
        geo = (ctypes.c_double * 6)(*utils.GeoTransform.from_bbox(stripe_bbox, h, w))
        s = sentinel(local.dtype, no_data_value)
        holder, nodata_ptr = _native.scalar_ptr(0 if s is None else s, local.dtype)
        thresholds = None
        if threshold_values is not None:
            thresholds = np.ascontiguousarray(threshold_values, dtype=np.float32)
        polys = soup.as_struct()
        desc = _frame_descriptor(local, 0)
        lib = _native.lib()
        _native.check(lib.gm_zonal_stats(
            ctypes.byref(desc), nodata_ptr, int(s is not None), ctypes.byref(polys), geo,
            _STAT_CODES["sum"], 0.0, None if thresholds is None else thresholds.ctypes.data, 0, h,
            None, covered.ctypes.data, partial.ctypes.data, _native.current_stream()))
    partial, covered = allreduce_partials(partial, covered, group)
    return finalize_partials(partial, statistic), np.nonzero(covered == 0)[0].tolist()


def _zonal_order_by_exchange(soup, local, no_data_value, stripe_bbox, has_rows, statistic, percentile,
                             threshold_values, group, with_covered=False):
    """Median / percentile of polygons that reach beyond a neighbouring stripe: every rank
    extracts the active values of its rows (gm_zonal_values), the segments are routed to the
    polygons' owner ranks ``p % world`` and selected there (gm_segment_order_stat)."""
    import ctypes

    from . import _native, utils
    from .geometry.aggregate import _STAT_CODES, _frame_descriptor
    from .raster._program import sentinel

    n = soup.n_polygons
    partial = np.zeros(n, dtype=PARTIAL_DTYPE)
    covered = np.zeros(n, dtype=np.int64)
    counts = values = None
    if has_rows and n > 0:
        _, h, w = local.shape
        geo = (ctypes.c_double * 6)(*utils.GeoTransform.from_bbox(stripe_bbox, h, w))
        s = sentinel(local.dtype, no_data_value)
        holder, nodata_ptr = _native.scalar_ptr(0 if s is None else s, local.dtype)
        thresholds = None
        if threshold_values is not None:
            thresholds = np.ascontiguousarray(threshold_values, dtype=np.float32)
        polys = soup.as_struct()
        desc = _frame_descriptor(local, 0)
        _native.check(_native.lib().gm_zonal_stats(
            ctypes.byref(desc), nodata_ptr, int(s is not None), ctypes.byref(polys), geo,
            _STAT_CODES["sum"], 0.0, None if thresholds is None else thresholds.ctypes.data, 0, h,
            None, covered.ctypes.data, partial.ctypes.data, _native.current_stream()))
        counts, values = _native.zonal_values(desc, nodata_ptr, int(s is not None), polys, geo,
                                              thresholds, partial["count"])
    if counts is None:
        counts = np.zeros(n, dtype=np.int64)
        values = np.zeros(0, dtype=local.dtype)
    owned, offsets, merged = exchange_segments(counts, values, group)
    mine = segment_order_statistic(merged, offsets, statistic, percentile)
    result = _gather_owned(mine, owned, n, group)
    if not with_covered:
        return result
    import torch

    dist = _dist()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    total = torch.from_numpy(covered).to(dev)
    if _world(group)[1] > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return result, total.cpu().numpy()


def _polygon_rows(soup, bbox, height):
    """(top, bottom) raster rows of the full grid that every polygon can touch, one row of
    slack either side; polygons without vertices get (height, -1).  Cached on the soup."""
    key = (tuple(bbox), int(height))
    cached = getattr(soup, "_row_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    n = soup.n_polygons
    y1, y2 = bbox[1], bbox[3]
    dy = (y2 - y1) / height
    starts = soup.ring_offsets[soup.poly_offsets[:-1]]
    ends = soup.ring_offsets[soup.poly_offsets[1:]]
    top = np.full(n, height, dtype=np.int64)
    bottom = np.full(n, -1, dtype=np.int64)
    filled = np.nonzero(ends > starts)[0]
    if len(filled):
        py = (y2 - soup.xy[:, 1]) / dy
        top[filled] = np.floor(np.minimum.reduceat(py, starts[filled])).astype(np.int64) - 1
        bottom[filled] = np.floor(np.maximum.reduceat(py, starts[filled])).astype(np.int64) + 1
    soup._row_cache = (key, (top, bottom))
    return top, bottom


def _stripe_ownership(soup, bbox, height, world):
    """Which rank selects which polygon when the raster is cut in `world` row stripes (cached
    on the soup): owner = the stripe holding the polygon's first row; `near` polygons cross
    into the next stripe only, `far` ones reach beyond it; per rank the rows it needs from
    its southern neighbour (`halo`) and how many of its own last rows its crossers span."""
    key = (tuple(bbox), int(height), int(world))
    cached = getattr(soup, "_owner_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    stops = np.array([b[1] for b in stripe_rows(height, world)], dtype=np.int64)
    top, bottom = _polygon_rows(soup, bbox, height)
    top_c, bottom_c = np.clip(top, 0, height - 1), np.clip(bottom, 0, height - 1)
    outside = (bottom < 0) | (top > height - 1)
    owner = np.searchsorted(stops, top_c, side="right")
    owner[outside] = 0
    crosses = ~outside & (bottom_c >= stops[owner])
    next_stop = stops[np.minimum(owner + 1, world - 1)]
    far = crosses & ((owner == world - 1) | (bottom_c >= next_stop))
    near = crosses & ~far
    halo = np.zeros(world, dtype=np.int64)
    keep = np.zeros(world, dtype=np.int64)
    if near.any():
        np.maximum.at(halo, owner[near], bottom_c[near] + 1 - stops[owner[near]])
        np.maximum.at(keep, owner[near], stops[owner[near]] - top_c[near])
    result = (owner, near, far, crosses, halo, keep)
    soup._owner_cache = (key, result)
    return result


def _as_tensor(local):
    """numpy / DeviceArray / torch tensor -> torch tensor sharing the memory."""
    import torch

    from . import _native

    if hasattr(local, "data_ptr"):
        return local
    if _native.is_device(local):
        owner = local
        while isinstance(owner, _native.DeviceArray):
            if hasattr(owner._owner, "data_ptr"):
                return owner._owner.reshape(local.shape)
            if owner._owner is None:
                break
            owner = owner._owner

        class _View(object):
            __cuda_array_interface__ = {"shape": local.shape, "typestr": local.dtype.str,
                                        "data": (local.ptr, False), "version": 2}

        view = torch.as_tensor(_View(), device="cuda")
        view._keepalive = local
        return view
    return torch.from_numpy(np.ascontiguousarray(local))


def _order_stat_call(soup, payload, no_data_value, bbox, threshold_values, statistic, percentile):
    """gm_zonal_stats median / percentile of ALL polygons over one raster: (out, covered)."""
    import ctypes

    from . import _native, utils
    from .geometry.aggregate import _STAT_CODES, _frame_descriptor
    from .raster._program import sentinel

    n = soup.n_polygons
    _, h, w = payload.shape
    geo = (ctypes.c_double * 6)(*utils.GeoTransform.from_bbox(bbox, h, w))
    s = sentinel(payload.dtype, no_data_value)
    holder, nodata_ptr = _native.scalar_ptr(0 if s is None else s, payload.dtype)
    thresholds = None
    if threshold_values is not None:
        thresholds = np.ascontiguousarray(threshold_values, dtype=np.float32)
    polys = soup.as_struct()
    if not _native.is_device(payload):
        payload = np.ascontiguousarray(payload)
    desc = _frame_descriptor(payload, 0)
    out = _native.pinned_empty((n,), np.float32)
    covered = _native.pinned_empty((n,), np.int64)
    covered[:] = 0
    _native.check(_native.lib().gm_zonal_stats(
        ctypes.byref(desc), nodata_ptr, int(s is not None), ctypes.byref(polys), geo,
        _STAT_CODES[statistic], float(percentile or 0.0),
        None if thresholds is None else thresholds.ctypes.data, 0, h,
        out.ctypes.data, covered.ctypes.data, None, _native.current_stream()))
    return out, covered


_HELPERS = {}


def _helper_pool():
    """One helper thread per process for the boundary work of the striped order statistics."""
    pool = _HELPERS.get("pool")
    if pool is None:
        from concurrent.futures import ThreadPoolExecutor

        pool = _HELPERS["pool"] = ThreadPoolExecutor(max_workers=1, thread_name_prefix="gm-boundary")
    return pool


def _side_stream(device):
    import torch

    key = ("stream", device)
    if key not in _HELPERS:
        _HELPERS[key] = torch.cuda.Stream(device=device)
    return _HELPERS[key]


def _rank_soups(soup, bbox, height, world, rank):
    """This rank's share of the order statistics, cached on the soup: the polygons it owns that
    lie inside its stripe and the ones that cross into the next stripe, each as ids + a soup of
    just those polygons (kept resident in HBM when a device is in use), so that a call visits
    ~N / world polygons instead of all N."""
    key = (tuple(bbox), int(height), int(world), int(rank))
    cache = getattr(soup, "_rank_soups", None)
    if cache is None:
        cache = soup._rank_soups = {}
    hit = cache.get(key)
    if hit is None:
        owner, near, far, crosses, _, _ = _stripe_ownership(soup, bbox, height, world)
        top, bottom = _polygon_rows(soup, bbox, height)
        outside = (bottom < 0) | (top > height - 1)
        inside_ids = np.nonzero((owner == rank) & ~crosses & ~outside)[0]
        near_ids = np.nonzero((owner == rank) & near)[0]
        hit = cache[key] = (inside_ids, soup.subset(inside_ids), near_ids, soup.subset(near_ids))
    return hit


def _zonal_order_striped(soup, local, no_data_value, bbox, height, rows, statistic, percentile,
                         threshold_values, group):
    """Median / percentile over a raster sharded in row stripes.

    Every polygon is OWNED by the rank whose stripe holds its first row.  A polygon that lies
    inside one stripe is selected there by the single-GPU kernel; one that crosses into the
    next stripe is selected by its owner on a boundary strip = the owner's last rows + the
    rows it needs from its southern neighbour (one ncclSend/ncclRecv of a few hundred rows).
    Each rank only visits the polygons it owns (`_rank_soups`).  The owners' results and
    covered-cell counts travel in ONE all-gather.  Only polygons that reach beyond the
    neighbouring stripe go through the value exchange (`_zonal_order_by_exchange`)."""
    import torch

    from . import _native

    rank, world = _world(group)
    dist = _dist()
    n = soup.n_polygons
    r0, r1 = rows
    x1, y1, x2, y2 = bbox
    dy = (y2 - y1) / height
    owner, near, far, crosses, halo, keep = _stripe_ownership(soup, bbox, height, world)
    inside_ids, inside_soup, near_ids, near_soup = _rank_soups(soup, bbox, height, world, rank)
    on_device = _native.is_device(local) or hasattr(local, "data_ptr")
    if on_device:
        inside_soup.to_device()
        near_soup.to_device()
    tensor = _as_tensor(local) if r1 > r0 else None
    stripe_bbox = (x1, y2 - r1 * dy, x2, y2 - r0 * dy)
    # owners publish their results in one all-gather of a (2, N) float32 block: row 0 the
    # statistic of the polygons this rank owns (others stay 0), row 1 whether they cover any cell
    # (all the caller needs of the counts: polygons without cells take the centroid fallback).
    # Under NCCL the owners' entries are picked on the device, so only (2, N) floats come back to
    # the host however many ranks there are.  Page-locked staging buffers and the owner -> slot
    # index are kept with the soup.
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    key = (tuple(bbox), int(height), int(world), dev)
    buffers = getattr(soup, "_gather_buffers", None)
    if buffers is None or buffers[0] != key:
        top, bottom = _polygon_rows(soup, bbox, height)
        stage_in = torch.empty((2, n), dtype=torch.float32)
        stage_out = torch.empty((2, n), dtype=torch.float32)
        if dev == "cuda":
            stage_in, stage_out = stage_in.pin_memory(), stage_out.pin_memory()
        pick = torch.from_numpy(owner.astype(np.int64) * (2 * n) + np.arange(n))
        pick = torch.stack([pick, pick + n]).to(dev)
        buffers = soup._gather_buffers = (key, stage_in, stage_out, pick, (bottom < 0) | (top > height - 1))
    _, stage_in, stage_out, pick, nowhere = buffers
    mine = stage_in.numpy()
    mine[...] = 0

    def thresholds_of(ids):
        return None if threshold_values is None else np.asarray(threshold_values)[ids]

    # boundary rows: my first rows go north, the southern neighbour's first rows come here, and the
    # polygons that cross into the next stripe are selected on the strip made of both.  Under NCCL
    # all of that runs on a helper thread with its own stream WHILE this thread selects the inner
    # polygons: the exchange, the few hundred boundary polygons and their three launches hide
    # behind the stripe's own select (the library calls release the GIL).
    link = "cuda" if dist.get_backend(group) == "nccl" else "cpu"

    def boundary(ready=None, device=None):
        if device is not None:           # helper thread: its own device context and stream
            torch.cuda.set_device(device)
            side = _side_stream(device)
            side.wait_event(ready)       # the stripe as the caller's stream left it
            with torch.cuda.stream(side), _native.use_stream(side.cuda_stream):
                return boundary()
        ops, recv = [], None
        if rank > 0 and halo[rank - 1] > 0:
            send = tensor[:, :int(halo[rank - 1])].to(link).contiguous()
            ops.append(dist.P2POp(dist.isend, send, rank - 1, group))
        if rank < world - 1 and halo[rank] > 0:
            recv = torch.empty((1, int(halo[rank]), tensor.shape[2]), dtype=tensor.dtype, device=link)
            ops.append(dist.P2POp(dist.irecv, recv, rank + 1, group))
        for work in (dist.batch_isend_irecv(ops) if ops else []):
            work.wait()
        if recv is None or not len(near_ids):
            if link == "cuda" and ops:
                torch.cuda.current_stream().synchronize()    # my rows have left before they may change
            return None
        k = int(keep[rank])
        strip = torch.cat([tensor[:, r1 - r0 - k:], recv.to(tensor.device)], dim=1).contiguous()
        strip_bbox = (x1, y2 - (r1 + int(halo[rank])) * dy, x2, y2 - (r1 - k) * dy)
        if strip.is_cuda:
            torch.cuda.current_stream().synchronize()   # the strip is complete before the library reads it
        return _order_stat_call(near_soup, _as_payload(strip), no_data_value, strip_bbox,
                                thresholds_of(near_ids), statistic, percentile)

    helper = None
    if link == "cuda" and tensor is not None and tensor.is_cuda:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        if _native.current_stream() not in (None, torch.cuda.current_stream().cuda_stream):
            _native.synchronize()      # the library's stream holds work on the stripe: let it finish
        helper = _helper_pool().submit(boundary, ready, tensor.device.index)
    _trace("boundary rows posted")
    try:
        if r1 > r0 and len(inside_ids):
            got, cov = _order_stat_call(inside_soup, local, no_data_value, stripe_bbox, thresholds_of(inside_ids),
                                        statistic, percentile)
            mine[0, inside_ids], mine[1, inside_ids] = got, cov > 0
    except BaseException:
        if helper is not None:      # never leave the helper's send / recv behind a raised error
            try:
                helper.result()
            except BaseException:
                pass
        raise
    _trace("select inside the stripe")
    strip_result = helper.result() if helper is not None else boundary()
    if strip_result is not None:
        got, cov = strip_result
        mine[0, near_ids], mine[1, near_ids] = got, cov > 0
    _trace("select on the boundary strip")
    block = stage_in.to(dev, non_blocking=True)
    gathered = torch.empty((world,) + tuple(block.shape), dtype=block.dtype, device=dev)
    if dev == "cuda":
        dist.all_gather_into_tensor(gathered, block, group=group)
        stage_out.copy_(gathered.reshape(-1)[pick], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        table = stage_out.numpy()
    else:
        dist.all_gather(list(gathered.unbind(0)), block, group=group)
        table = gathered.reshape(-1)[pick].numpy()
    _trace("all-gather + download")
    result = table[0].copy()
    covered = table[1].astype(np.int64)       # 0 / 1: only "covers no cell" is reported
    result[nowhere] = np.nan          # outside the raster: nobody selected them
    covered[nowhere] = 0
    if far.any():   # same on every rank: the collectives inside line up
        ids = np.nonzero(far)[0]
        sub = soup.subset(ids)
        result[ids], covered[ids] = _zonal_order_by_exchange(
            sub, local, no_data_value, stripe_bbox, r1 > r0, statistic, percentile, thresholds_of(ids), group,
            with_covered=True)
    return result, np.nonzero(covered == 0)[0].tolist()


def _zonal_striped_device(soup, local, no_data_value, stripe_bbox, statistic, threshold_values, group, has_rows):
    """count / sum / mean / min / max of a striped raster with everything but the final N
    results in HBM: stripe partials (gm_zonal_partials_device; only what the statistic needs),
    one SUM all-reduce (plus one MIN all-reduce for min / max) over NVLink, statistic + download
    (gm_zonal_finalize_device).  A 40000 x 40000 stripe pass takes a fraction of a millisecond, so
    everything that does not change between calls -- descriptors, partial vectors, page-locked
    result buffers -- is kept with the soup."""
    import ctypes

    import torch

    from . import _native, utils
    from .geometry.aggregate import _STAT_CODES, _frame_descriptor
    from .raster._program import sentinel

    dist = _dist()
    n = soup.n_polygons
    lib = _native.lib()
    stream = _native.current_stream()
    torch_stream = torch.cuda.current_stream().cuda_stream
    same_stream = stream is not None and int(stream) == int(torch_stream)
    shape = tuple(local.shape) if has_rows else None
    key = (tuple(stripe_bbox), shape, str(local.dtype) if has_rows else None, repr(no_data_value), n)
    kept = getattr(soup, "_stripe_call", None)
    if kept is None or kept["key"] != key:
        kept = {"key": key,
                "sums": torch.empty(3 * n, dtype=torch.float64, device="cuda"),
                "extremes": torch.empty(2 * n, dtype=torch.float64, device="cuda"),
                "out": _native.pinned_empty((n,), np.float32),
                "covered": _native.pinned_empty((n,), np.int64)}
        if has_rows:
            s = sentinel(local.dtype, no_data_value)
            kept["geo"] = (ctypes.c_double * 6)(*utils.GeoTransform.from_bbox(stripe_bbox, shape[1], shape[2]))
            kept["nodata"] = _native.scalar_ptr(0 if s is None else s, local.dtype)
            kept["has_nodata"] = int(s is not None)
            # resident soup: the stripe pass keeps its preparation (pixel-space vertices, row
            # ranges, the ids of the polygons with rows here) and does not synchronise
            kept["polys"] = soup.to_device().as_struct()
        soup._stripe_call = kept
    sums, extremes = kept["sums"], kept["extremes"]
    if not has_rows:      # (a stripe pass writes every entry of both vectors)
        sums.zero_()
        if statistic in ("min", "max"):
            extremes.fill_(float(np.finfo(np.float64).max))
    if has_rows:
        thresholds = None
        if threshold_values is not None:
            thresholds = np.ascontiguousarray(threshold_values, dtype=np.float32)
        desc = _frame_descriptor(local, 0)
        if not same_stream:
            torch.cuda.current_stream().synchronize()   # the tensors above are ready
        _native.check(lib.gm_zonal_partials_device(
            ctypes.byref(desc), kept["nodata"][1], kept["has_nodata"], ctypes.byref(kept["polys"]), kept["geo"],
            None if thresholds is None else thresholds.ctypes.data, 0, shape[1],
            sums.data_ptr(), extremes.data_ptr(), _STAT_CODES[statistic], stream))
        if not same_stream:
            _native.synchronize()
    _trace("stripe partials")
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    if statistic in ("min", "max"):       # the extremes only travel when they are asked for
        dist.all_reduce(extremes, op=dist.ReduceOp.MIN, group=group)
    _trace("all-reduce")
    if not same_stream:
        torch.cuda.current_stream().synchronize()
    # (the result buffers are reused by the next call on this soup: copy what must outlive it)
    out, covered = kept["out"], kept["covered"]
    _native.check(lib.gm_zonal_finalize_device(sums.data_ptr(), extremes.data_ptr(), n, _STAT_CODES[statistic],
                                               out.ctypes.data, covered.ctypes.data,
                                               ctypes.byref(kept["polys"]) if has_rows else None, stream))
    _trace("finalise + download")
    return out.copy(), np.nonzero(covered == 0)[0].tolist()


# ---------------------------------------------------------------------------
# order statistics: route every polygon's values to its owner
# ---------------------------------------------------------------------------


def exchange_segments(counts, values, group=None):
    """All-to-all-v of per-polygon value segments.

    `values` holds this rank's active cell values packed polygon by polygon
    (`counts[p]` values for polygon p).  Polygon p is owned by rank ``p % world``.
    Returns (owned polygon ids, segment offsets, values) for this rank, where the
    values of an owned polygon are the concatenation over all ranks."""
    import torch

    rank, world = _world(group)
    counts = np.asarray(counts, dtype=np.int64)
    values = np.ascontiguousarray(values)
    n = len(counts)
    starts = np.concatenate([[0], np.cumsum(counts)])
    owned = np.arange(rank, n, world)
    if world == 1:
        return owned, starts.copy(), values
    dist = _dist()
    # every rank learns every rank's counts (N int64 each)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    all_counts = [torch.empty(n, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_counts, torch.from_numpy(counts).to(dev), group=group)
    all_counts = np.stack([c.cpu().numpy() for c in all_counts])  # (world, N)
    # pack what goes to each destination: its polygons in ascending order
    send_chunks, send_sizes = [], []
    for dest in range(world):
        ids = np.arange(dest, n, world)
        chunk = [values[starts[p]:starts[p + 1]] for p in ids]
        chunk = np.concatenate(chunk) if chunk else values[:0]
        send_chunks.append(chunk)
        send_sizes.append(len(chunk))
    recv_sizes = [int(all_counts[src, owned].sum()) for src in range(world)]
    send = torch.from_numpy(np.concatenate(send_chunks) if send_chunks else values[:0]).to(dev)
    recv = torch.empty(sum(recv_sizes), dtype=send.dtype, device=dev)
    if dev == "cuda":
        dist.all_to_all_single(recv, send, recv_sizes, send_sizes, group=group)
    else:  # gloo has no all_to_all_single for uneven splits on every build: grouped send/recv
        ops, pieces, at = [], [], 0
        for src in range(world):
            pieces.append(recv[at:at + recv_sizes[src]])
            at += recv_sizes[src]
        at = 0
        for dest in range(world):
            piece = send[at:at + send_sizes[dest]]
            at += send_sizes[dest]
            if dest == rank:
                pieces[rank].copy_(piece)
                continue
            if send_sizes[dest]:
                ops.append(dist.P2POp(dist.isend, piece.contiguous(), dest, group))
        for src in range(world):
            if src != rank and recv_sizes[src]:
                ops.append(dist.P2POp(dist.irecv, pieces[src], src, group))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
    recv = recv.cpu().numpy()
    # received layout: for each source rank, its segments of my polygons in ascending id order;
    # regroup per polygon
    per_src_starts, at = [], 0
    for src in range(world):
        c = all_counts[src, owned]
        per_src_starts.append(at + np.concatenate([[0], np.cumsum(c)]))
        at += int(c.sum())
    totals = all_counts[:, owned].sum(axis=0)
    offsets = np.concatenate([[0], np.cumsum(totals)]).astype(np.int64)
    merged = np.empty(int(offsets[-1]), dtype=values.dtype)
    for k in range(len(owned)):
        at = offsets[k]
        for src in range(world):
            a, b = per_src_starts[src][k], per_src_starts[src][k + 1]
            merged[at:at + (b - a)] = recv[a:b]
            at += b - a
    return owned, offsets, merged


def segment_order_statistic(values, offsets, statistic, percentile=None):
    """Median / percentile of every segment ``values[offsets[k]:offsets[k+1]]`` on the GPU
    (radix select of libgeokernels.so; the interpolation arithmetic of
    measurements.py:132-137 and scipy.ndimage's median)."""
    from . import _native

    return _native.segment_order_statistic(np.ascontiguousarray(values), np.asarray(offsets, dtype=np.int64),
                                           statistic, percentile)


def _gather_owned(mine, owned, n, group=None):
    """Every rank contributes the results of its own polygons; all ranks get all N."""
    import torch

    rank, world = _world(group)
    out = np.full(n, np.nan, dtype=np.float32)
    out[owned] = mine
    if world == 1:
        return out
    dist = _dist()
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    # NaN marks "not mine": a max-reduce over ranks would lose NaN results, so gather instead
    gathered = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(out).to(dev), group=group)
    for src in range(world):
        ids = np.arange(src, n, world)
        out[ids] = gathered[src].cpu().numpy()[ids]
    return out
