"""Blocks that combine rasters into a larger one, on the CUDA evaluator.

Drop-in for the reference's raster/combine.py (``Group``, :140-464): rasters are merged along
x, y and time; where several hold data at the same instant the rightmost one shows, 'no data'
is transparent.  The planning (which source answers which frames, :199-298) is host Python as in
the reference; the pixel work -- the reference's loops of ``target[index] = source[index]`` per
frame (:316-343, :371-387) -- is one 'last' reduction program per output frame
(raster/reduction.py), i.e. one pass over the frames that contribute to it.
"""
import itertools
from datetime import timedelta as Timedelta

import numpy as np

from .. import _native
from ..utils import Extent, GeoTransform, get_dtype_max
from .base import RasterBlock
from .reduction import reduce_rasters

__all__ = ["Group"]


def _present(items):
    return [x for x in items if x is not None]


def _band(values, index):
    """Frame ``index`` of a (t, h, w) host or device array as a (1, h, w) view."""
    if _native.is_device(values):
        t, h, w = values.shape
        return _native.DeviceArray((1, h, w), values.dtype, owner=values,
                                   ptr=values.ptr + index * h * w * values.dtype.itemsize)
    return values[index:index + 1]


def _stack_frames(frames, shape, dtype):
    """(1, h, w) frames -> one (n, h, w) array (device when all frames are)."""
    if len(frames) == 1:
        return frames[0]
    if all(_native.is_device(f) for f in frames):
        out = _native.DeviceArray(shape, dtype)
        plane = shape[1] * shape[2] * np.dtype(dtype).itemsize
        lib = _native.lib()
        for i, f in enumerate(frames):
            _native.check(lib.gm_memcpy_d2d(out.ptr + i * plane, f.ptr, plane, _native.current_stream()))
        return out
    return np.concatenate([np.asarray(f) for f in frames], axis=0)


def _overlay_frames(contributions, shape, dtype):
    """``contributions[k]`` = [(values, frame index, no data value), ...] for output frame k, left
    to right; frames without a contribution hold the fill value."""
    fill = get_dtype_max(dtype)
    frames = []
    for parts in contributions:
        if not parts:
            frames.append(np.full((1,) + tuple(shape[1:]), fill, dtype=dtype))
            continue
        stack = [{"values": _band(v, i), "no_data_value": nd} for v, i, nd in parts]
        frames.append(reduce_rasters(stack, "last", fill, dtype)["values"])
    return {"values": _stack_frames(frames, shape, dtype), "no_data_value": fill}


class BaseCombine(RasterBlock):
    """Base of blocks that merge rasters: period and extent are the unions of the sources',
    the time step is kept when the sources are equidistant and aligned."""

    def __init__(self, *args):
        for arg in args:
            if not isinstance(arg, RasterBlock):
                raise TypeError("'{}' object is not allowed".format(type(arg)))
        super(BaseCombine, self).__init__(*args)

    @staticmethod
    def get_aligned_timedelta(sources):
        """The common time step of the sources that hold data, None when they differ or are not
        an integer number of steps apart."""
        pairs = [(s.timedelta, s.period) for s in sources]
        pairs = [(d, p) for d, p in pairs if d is not None and p is not None]
        if not pairs:
            return None
        delta = pairs[0][0]
        if any(d != delta for d, _ in pairs[1:]):
            return None
        seconds = delta.total_seconds()
        first = pairs[0][1][0]
        for _, (begin, _end) in pairs[1:]:
            if (first - begin).total_seconds() % seconds != 0:
                return None
        return delta

    @property
    def timedelta(self):
        return self.get_aligned_timedelta(self.args)

    @property
    def temporal(self):
        return any(x.temporal for x in self.args)

    @property
    def period(self):
        periods = _present(x.period for x in self.args)
        if not periods:
            return None
        return min(p[0] for p in periods), max(p[1] for p in periods)

    @property
    def extent(self):
        extents = _present(x.extent for x in self.args)
        if not extents:
            return None
        return (min(e[0] for e in extents), min(e[1] for e in extents),
                max(e[2] for e in extents), max(e[3] for e in extents))

    @property
    def dtype(self):
        return np.result_type(*self.args)

    @property
    def fillvalue(self):
        return get_dtype_max(self.dtype)

    @property
    def geometry(self):
        geometries = _present(x.geometry for x in self.args)
        if not geometries:
            return None
        boxes = [Extent.from_geometry(g).bbox for g in geometries]
        union = (min(b[0] for b in boxes), min(b[1] for b in boxes),
                 max(b[2] for b in boxes), max(b[3] for b in boxes))
        return Extent(union, getattr(geometries[0], "projection", None)).as_geometry()

    @property
    def projection(self):
        projection = self.args[0].projection
        if projection is None or any(arg.projection != projection for arg in self.args[1:]):
            return None
        return projection

    @property
    def geo_transform(self):
        first = self.args[0].geo_transform
        if first is None:
            return None
        first = GeoTransform(first)
        for arg in self.args[1:]:
            other = arg.geo_transform
            if other is None or not first.aligns_with(other):
                return None
        return first


class Group(BaseCombine):
    """Combine rasters along x, y and time.  Values of rasters further to the right show on
    top; 'no data' is transparent (reference: raster/combine.py:140-464)."""

    def get_relevant_sources(self, start, stop):
        stores = [s for s in self.args if s.period is not None]
        if not stores:
            return []
        begins, ends = zip(*(s.period for s in stores))
        if start is None:        # the latest frame: the store(s) that end last
            last = max(ends)
            return [s for e, s in zip(ends, stores) if e == last]
        if stop is None:         # one instant: the stores containing it, else the nearest
            inside = [s for b, e, s in zip(begins, ends, stores) if b <= start <= e]
            if inside:
                return inside
            nearest = min(begins + ends, key=lambda d: abs(d - start))
            return [s for d, s in zip(ends + begins, stores + stores) if d == nearest]
        return [s for b, e, s in zip(begins, ends, stores) if not (stop < b or start > e)]

    def get_sources_and_requests(self, **request):
        start, stop, mode = request.get("start"), request.get("stop"), request["mode"]
        nothing = [(dict(combine_mode="simple"), None)]
        period = self.period
        if period is None:
            return nothing
        if start is not None and stop is not None and (start > period[1] or stop < period[0]):
            return nothing
        timedelta = self.timedelta
        if timedelta is None:
            # no common time axis: merge on the timestamps the sources report
            sources = self.get_relevant_sources(start, stop)
            if not sources:
                return nothing
            requests = [(s, request) for s in sources]
            if mode != "time":
                requests += [(s, dict(mode="time", start=start, stop=stop)) for s in sources]
            process_kwargs = dict(combine_mode="by_time", mode=mode, start=start, stop=stop)
        else:
            # a common, aligned time axis: every source fills a slice of result frames
            step = timedelta.total_seconds()
            origin = period[0]
            if start is None:
                start = period[1]
            elif start < period[0]:
                start = period[0]
            else:   # up to the next frame
                start += Timedelta(seconds=(origin - start).total_seconds() % step)
            if stop is None:
                stop = start
            elif stop > period[1]:
                stop = period[1]
            else:   # down to the previous frame
                stop -= Timedelta(seconds=(stop - origin).total_seconds() % step)
            if mode == "time":
                return [(dict(combine_mode="by_bands", mode=mode, start=start, stop=stop,
                              timedelta=timedelta), None)]
            requests, bands = [], []
            for source in self.get_relevant_sources(start, stop):
                first, last = max(start, source.period[0]), min(stop, source.period[1])
                bands.append((int((first - start).total_seconds() // step),
                              int((last - start).total_seconds() // step) + 1))
                requests.append((source, dict(request, start=first, stop=last)))
            process_kwargs = dict(combine_mode="by_bands", mode=mode, bands=bands)
            n_frames = int((stop - start).total_seconds() // step) + 1
            if mode == "meta":
                process_kwargs["nbands"] = n_frames
            if mode == "vals":
                process_kwargs["shape"] = (n_frames, request["height"], request["width"])
        if mode == "vals":
            process_kwargs["dtype"] = self.dtype
        return [(process_kwargs, None)] + requests

    @staticmethod
    def _unique_times(multi):
        return sorted(set(itertools.chain(*_present(d.get("time") for d in multi))))

    @staticmethod
    def _nearest_index(times, start):
        if start is None:
            return len(times) - 1
        return min(range(len(times)), key=lambda i: abs(times[i] - start))

    @staticmethod
    def _merge_vals_by_time(multi, times, kwargs):
        instants = Group._unique_times(times)
        frame_of = {t: k for k, t in enumerate(instants)}
        dtype = np.dtype(kwargs["dtype"])
        contributions = [[] for _ in instants]
        for data, time in zip(multi, times):
            for index, instant in enumerate(time["time"]):
                contributions[frame_of[instant]].append((data["values"], index, data["no_data_value"]))
        shape = (len(instants),) + tuple(multi[0]["values"].shape[1:])
        if kwargs["stop"] is None and len(instants) > 1:    # a single frame is wanted
            k = Group._nearest_index(instants, kwargs["start"])
            contributions, shape = contributions[k:k + 1], (1,) + shape[1:]
        return _overlay_frames(contributions, shape, dtype)

    @staticmethod
    def _merge_meta_by_time(multi, times, kwargs):
        instants = Group._unique_times(times)
        frame_of = {t: k for k, t in enumerate(instants)}
        result = [None] * len(instants)
        for data, time in zip(multi, times):
            for index, instant in enumerate(time["time"]):
                result[frame_of[instant]] = data["meta"][index]
        if kwargs["stop"] is None and len(instants) > 1:
            k = Group._nearest_index(instants, kwargs["start"])
            result = result[k:k + 1]
        return {"meta": result}

    @staticmethod
    def _merge_vals_by_bands(multi, bands, dtype, shape):
        contributions = [[] for _ in range(shape[0])]
        for data, (a, b) in zip(multi, bands):
            for k in range(a, b):
                contributions[k].append((data["values"], k - a, data["no_data_value"]))
        return _overlay_frames(contributions, tuple(shape), np.dtype(dtype))

    @staticmethod
    def _merge_meta_by_bands(multi, bands, nbands):
        result = [""] * nbands
        for data, (a, b) in zip(multi, bands):
            for k, meta in zip(range(a, b), data["meta"]):
                if meta:
                    result[k] = meta
        return {"meta": result}

    @staticmethod
    def process(process_kwargs, *args):
        combine_mode, mode = process_kwargs["combine_mode"], process_kwargs.get("mode")
        if combine_mode == "simple":
            return None
        if combine_mode == "by_time" and mode == "time":
            instants = Group._unique_times(args)
            if process_kwargs["stop"] is None and len(instants) > 1:
                k = Group._nearest_index(instants, process_kwargs["start"])
                instants = instants[k:k + 1]
            return {"time": instants}
        if combine_mode == "by_time" and mode in ("meta", "vals"):
            half = len(args) // 2       # payloads first, then the matching time answers
            multi, times = _present(args[:half]), _present(args[half:])
            if not multi:
                return None
            if mode == "vals":
                return Group._merge_vals_by_time(multi, times, process_kwargs)
            return Group._merge_meta_by_time(multi, times, process_kwargs)
        if combine_mode == "by_bands" and mode == "time":
            start, stop, delta = process_kwargs["start"], process_kwargs["stop"], process_kwargs["timedelta"]
            count = int((stop - start).total_seconds() // delta.total_seconds()) + 1
            return {"time": [start + i * delta for i in range(count)]}
        if combine_mode == "by_bands" and mode in ("meta", "vals"):
            pairs = [(d, b) for d, b in zip(args, process_kwargs["bands"]) if d is not None]
            multi, bands = [d for d, _ in pairs], [b for _, b in pairs]
            if mode == "vals":
                return Group._merge_vals_by_bands(multi, bands, process_kwargs["dtype"], process_kwargs["shape"])
            return Group._merge_meta_by_bands(multi, bands, process_kwargs["nbands"])
        raise ValueError("Unknown combine_mode / mode combination")
