"""TemporalAggregate and Cumulative on the GPU.

Drop-in for the two aggregating blocks of the reference's raster/temporal.py
(:480-768 and :775-1005).  The time-axis bookkeeping (which frames fall in
which pandas resample bin) stays on the host, as in the reference; the pixel
work -- a reduction or running sum over the band axis -- is one CUDA pass
over the (T, H, W) stack (csrc/gm_temporal.cu).  Snap / Shift / Resample only
relabel the time axis and are outside the hot path.
"""
import ctypes
from datetime import timedelta as Timedelta

import numpy as np
import pandas as pd
from pandas.tseries.frequencies import to_offset

from .. import _native, _state
from ..utils import dtype_for_statistic, get_dtype_max, parse_percentile_statistic
from .base import BaseSingle, RasterBlock

__all__ = ["TemporalAggregate", "Cumulative"]

# pandas closes/labels these frequencies on the right (TimeGrouper.__init__)
_END_ANCHORED = {"M", "A", "Q", "BM", "BA", "BQ", "W", "ME", "YE", "QE", "BME", "BYE", "BQE"}
_RENAMED = {  # aliases removed in pandas 3 (reference utils.py:43-59)
    "M": "ME", "BM": "BME", "SM": "SME", "CBM": "CBME", "Q": "QE", "BQ": "BQE", "Y": "YE",
    "BY": "BYE", "A": "YE", "BA": "BYE", "AS": "YS", "BAS": "BYS", "H": "h", "BH": "bh",
    "CBH": "cbh", "T": "min", "S": "s", "L": "ms", "U": "us", "N": "ns",
}
MICROSECOND = Timedelta(microseconds=1)
_STAT_CODES = {"sum": 0, "count": 1, "min": 2, "max": 3, "mean": 4, "median": 5, "std": 6,
               "var": 7, "percentile": 8}


def normalize_offset(freq):
    """Frequency string in the spelling of the installed pandas (utils.py:979-1011)."""
    if freq is None:
        return None
    digits = freq[: len(freq) - len(freq.lstrip("0123456789"))]
    alias = freq[len(digits):]
    if alias in _RENAMED:
        try:
            to_offset(freq)
        except ValueError:
            freq = digits + _RENAMED[alias]
    return to_offset(freq).freqstr


def offset_to_timedelta(freq):
    """timedelta of a fixed frequency, None for calendar dependent ones (utils.py:960-976)."""
    try:
        return pd.Timedelta(to_offset(freq).nanos, unit="ns").to_pytimedelta()
    except (ValueError, AttributeError):
        return None


def _check_timezone(name):
    if not isinstance(name, str):
        raise TypeError("'{}' object is not allowed.".format(type(name)))
    try:
        import pytz

        return pytz.timezone(name).zone
    except ImportError:
        import zoneinfo

        if name.upper() == "UTC":
            return "UTC"
        try:
            return zoneinfo.ZoneInfo(name).key
        except Exception:
            raise ValueError("Unknown time zone '{}'".format(name))


def _dt_to_ts(dt, timezone):
    return pd.Timestamp(dt, tz="UTC").tz_convert(timezone)


def _ts_to_dt(ts, timezone):
    if ts.tzinfo is None:
        ts = ts.tz_localize(timezone)
    return ts.tz_convert("UTC").tz_localize(None).to_pydatetime(warn=False)


def _get_bin_label(dt, frequency, closed, label, timezone):
    """Label of the resample bin that holds ``dt`` (raster/temporal.py:272-284)."""
    series = pd.Series([0], index=[_dt_to_ts(dt, timezone)])
    found = None
    for found, members in series.resample(frequency, closed=closed, label=label):
        if len(members):
            break
    return _ts_to_dt(found, timezone)


def _get_bin_start(dt, frequency, closed, label, timezone):
    series = pd.Series([0], index=[_dt_to_ts(dt, timezone)])
    return series.resample(frequency, closed=closed, label="left").first().index[0]


def _get_closest_label(dt, frequency, timezone, side="both"):
    """Nearest bin label to ``dt`` (raster/temporal.py:322-349)."""
    ts = _dt_to_ts(dt, timezone)
    centre = _dt_to_ts(_get_bin_label(dt, frequency, "left", "left", timezone), timezone)
    freq = to_offset(frequency)
    candidates = pd.date_range(centre - freq, centre + freq, freq=freq)
    delta = (candidates - ts).to_series()
    delta.index = candidates
    if side == "right":
        delta = delta[delta >= pd.Timedelta(0)]
    elif side == "left":
        delta = delta[delta <= pd.Timedelta(0)]
    return _ts_to_dt(delta.abs().idxmin(), timezone)


def _default_closed_label(frequency, closed, label):
    if frequency is None:
        return "right", "right"
    rule = to_offset(frequency).rule_code
    side = "right" if rule.split("-")[0] in _END_ANCHORED else "left"
    return closed or side, label or side


def _label_to_bin_start(dt, frequency, closed, label, timezone):
    ts = _dt_to_ts(dt, timezone)
    if label == "right":
        ts -= to_offset(frequency)
    if closed == "right":
        ts += MICROSECOND
    return _ts_to_dt(ts, timezone)


def _label_to_bin_end(dt, frequency, closed, label, timezone):
    ts = _dt_to_ts(dt, timezone)
    if label == "left":
        ts += to_offset(frequency)
    if closed == "left":
        ts -= MICROSECOND
    return _ts_to_dt(ts, timezone)


def _resampled_period(period, frequency, closed, label, timezone):
    if period is None:
        return None
    if frequency is None:
        return period[-1], period[-1]
    return tuple(_get_bin_label(x, frequency, closed, label, timezone) for x in period)


def _snap_to_resampled_labels(period, start, stop, frequency, timezone):
    """(start label, stop label) of a request; (None, None) when nothing is in range
    (raster/temporal.py:404-450)."""
    if period is None:
        return None, None
    first, last = period
    if start is None:
        start = last
    if stop is None:
        if start <= first:
            return first, None
        if start >= last:
            return last, None
        return _get_closest_label(start, frequency, timezone, side="both"), None
    if start > last or stop < first:
        return None, None
    start = first if start <= first else _get_closest_label(start, frequency, timezone, side="right")
    stop = last if stop >= last else _get_closest_label(stop, frequency, timezone, side="left")
    if start > stop:
        return None, None
    return start, stop


def _bins_to_arrays(groups):
    """CSR (offsets, frame indices) int32 arrays of a list of frame index lists."""
    offsets = np.zeros(len(groups) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(g) for g in groups])
    frames = np.fromiter((i for g in groups for i in g), dtype=np.int32, count=int(offsets[-1]))
    if len(frames) == 0:
        frames = np.zeros(1, dtype=np.int32)
    return offsets, frames


def _nodata_arg(values, no_data_value):
    from ._program import sentinel

    s = sentinel(values.dtype, no_data_value)
    holder, ptr = _native.scalar_ptr(0 if s is None else s, values.dtype)
    return holder, ptr, int(s is not None)


def _run(values, out_shape, out_dtype, launch):
    lib = _native.lib()
    on_device = _native.is_device(values) or _state.keep_on_device()
    if on_device and not _native.is_device(values):
        values = _native.DeviceArray.from_host(values)
    if not on_device:
        values = np.ascontiguousarray(values)
    out = (_native.DeviceArray(out_shape, out_dtype) if on_device
           else _native.pinned_empty(out_shape, out_dtype))
    src, dst = _native.as_gm_array(values), _native.as_gm_array(out)
    _native.check(launch(lib, ctypes.byref(src), ctypes.byref(dst), _native.current_stream()))
    if on_device and not _state.keep_on_device():
        out = out.to_host()
    return out


def _resample_indices(times, frequency, closed, label, timezone):
    series = pd.Series(index=times, dtype=float).tz_localize("UTC").tz_convert(timezone)
    return series, series.resample(frequency, closed=closed, label=label).indices


class _StatisticMixin(object):
    @staticmethod
    def _parse(statistic, allowed):
        if not isinstance(statistic, str):
            raise TypeError("'{}' object is not allowed.".format(type(statistic)))
        name, q = parse_percentile_statistic(statistic.lower())
        if q:
            return "p{0}".format(q)
        if name not in allowed:
            raise ValueError("Unknown statistic '{}'".format(name))
        return name


class TemporalAggregate(_StatisticMixin, BaseSingle):
    """Aggregate frames per pandas resample bin (``frequency``) or over the whole
    period (``frequency=None``).  statistic: sum, count, min, max, mean, median,
    std, var or p<percentile> (reference: raster/temporal.py:480-768)."""

    STATISTICS = {
        "sum": {"extensive": True}, "count": {"extensive": True}, "min": {"extensive": False},
        "max": {"extensive": False}, "mean": {"extensive": False}, "median": {"extensive": False},
        "std": {"extensive": False}, "var": {"extensive": False},
    }

    def __init__(self, source, frequency, statistic="sum", closed=None, label=None, timezone="UTC"):
        if not isinstance(source, RasterBlock):
            raise TypeError("'{}' object is not allowed.".format(type(source)))
        if frequency is not None:
            if not isinstance(frequency, str):
                raise TypeError("'{}' object is not allowed.".format(type(frequency)))
            frequency = normalize_offset(frequency)
            if closed not in {None, "left", "right"}:
                raise ValueError("closed must be None, 'left', or 'right'.")
            if label not in {None, "left", "right"}:
                raise ValueError("label must be None, 'left', or 'right'.")
            timezone = _check_timezone(timezone)
        else:
            closed = label = timezone = None
        statistic = self._parse(statistic, self.STATISTICS)
        super(TemporalAggregate, self).__init__(source, frequency, statistic, closed, label, timezone)

    source = property(lambda self: self.args[0])
    statistic = property(lambda self: self.args[2])
    closed = property(lambda self: self.args[3])
    label = property(lambda self: self.args[4])
    timezone = property(lambda self: self.args[5])

    @property
    def frequency(self):
        return normalize_offset(self.args[1])

    @property
    def _snap_kwargs(self):
        closed, label = _default_closed_label(self.frequency, self.closed, self.label)
        return {"frequency": self.frequency, "closed": closed, "label": label, "timezone": self.timezone}

    @property
    def period(self):
        return _resampled_period(self.source.period, **self._snap_kwargs)

    @property
    def timedelta(self):
        return None if self.frequency is None else offset_to_timedelta(self.frequency)

    @property
    def temporal(self):
        return self.frequency is not None

    @property
    def dtype(self):
        return dtype_for_statistic(self.source.dtype, self.statistic)

    @property
    def fillvalue(self):
        return get_dtype_max(self.dtype)

    def get_sources_and_requests(self, **request):
        kwargs = self._snap_kwargs
        mode = request["mode"]
        start_label, stop_label = _snap_to_resampled_labels(
            self.period, request.get("start"), request.get("stop"),
            frequency=self.frequency, timezone=self.timezone,
        )
        if start_label is None:
            return [({"empty": True, "mode": mode}, None)]
        kwargs.update(mode=mode, start=start_label, stop=stop_label)
        if mode == "time":
            return [(kwargs, None)]
        if self.frequency is None:
            request["start"], request["stop"] = self.source.period
        else:
            request["start"] = _label_to_bin_start(
                start_label, kwargs["frequency"], kwargs["closed"], kwargs["label"], kwargs["timezone"])
            request["stop"] = _label_to_bin_end(
                stop_label or start_label, kwargs["frequency"], kwargs["closed"], kwargs["label"],
                kwargs["timezone"])
        if mode == "vals":
            kwargs["dtype"] = np.dtype(self.dtype).str
            kwargs["statistic"] = self.statistic
        time_request = {"mode": "time", "start": request["start"], "stop": request["stop"]}
        if "time_resolution" in request:
            time_request["time_resolution"] = request["time_resolution"]
        return [(kwargs, None), (self.source, time_request), (self.source, request)]

    @staticmethod
    def process(process_kwargs, time_data=None, data=None):
        mode = process_kwargs["mode"]
        empty = None if mode == "vals" else {mode: []}
        if process_kwargs.get("empty"):
            return empty
        start, stop = process_kwargs["start"], process_kwargs["stop"]
        frequency = process_kwargs["frequency"]
        if frequency is None:
            labels = pd.DatetimeIndex([start])
        else:
            labels = pd.date_range(start, stop or start, freq=frequency)
        if mode == "time":
            return {"time": labels.to_pydatetime().tolist()}
        if time_data is None or not time_data.get("time"):
            return empty

        timezone = process_kwargs["timezone"]
        times = time_data["time"]
        labels = labels.tz_localize("UTC").tz_convert(timezone)
        if frequency is None:
            indices = {labels[0]: range(len(times))}
        else:
            _, indices = _resample_indices(times, frequency, process_kwargs["closed"],
                                           process_kwargs["label"], timezone)
        if mode == "meta":
            if data is None or "meta" not in data:
                return {"meta": []}
            meta = data["meta"]
            return {"meta": [[meta[i] for i in indices[ts]] for ts in labels]}

        if data is None or "values" not in data:
            return None
        values = data["values"]
        if values.shape[0] != len(times):
            raise RuntimeError("Shape of raster does not match number of timestamps")
        statistic, q = parse_percentile_statistic(process_kwargs["statistic"])
        dtype = np.dtype(process_kwargs["dtype"])
        groups = [list(indices.get(ts, ())) for ts in labels]
        offsets, frames = _bins_to_arrays(groups)
        holder, nodata_ptr, has_nodata = _nodata_arg(values, data["no_data_value"])
        out = _run(
            values, (len(labels), values.shape[1], values.shape[2]), dtype,
            lambda lib, src, dst, stream: lib.gm_temporal_aggregate(
                src, dst, nodata_ptr, has_nodata, _STAT_CODES[statistic], float(q or 0.0),
                offsets.ctypes.data, frames.ctypes.data, len(groups), stream),
        )
        return {"values": out, "no_data_value": get_dtype_max(dtype)}


class Cumulative(_StatisticMixin, BaseSingle):
    """Running sum / count over time, restarted every ``frequency``
    (reference: raster/temporal.py:775-1005)."""

    STATISTICS = {"sum": {"extensive": True}, "count": {"extensive": True}}

    def __init__(self, source, statistic="sum", frequency=None, timezone="UTC"):
        if not isinstance(source, RasterBlock):
            raise TypeError("'{}' object is not allowed.".format(type(source)))
        statistic = self._parse(statistic, self.STATISTICS)
        if frequency is not None:
            if not isinstance(frequency, str):
                raise TypeError("'{}' object is not allowed.".format(type(frequency)))
            frequency = normalize_offset(frequency)
            timezone = _check_timezone(timezone)
        else:
            timezone = None
        super().__init__(source, statistic, frequency, timezone)

    source = property(lambda self: self.args[0])
    statistic = property(lambda self: self.args[1])
    timezone = property(lambda self: self.args[3])

    @property
    def frequency(self):
        return normalize_offset(self.args[2])

    @property
    def _snap_kwargs(self):
        return {"frequency": self.frequency, "closed": "right", "label": "right",
                "timezone": self.timezone}

    @property
    def dtype(self):
        return dtype_for_statistic(self.source.dtype, self.statistic)

    @property
    def fillvalue(self):
        return get_dtype_max(self.dtype)

    def get_sources_and_requests(self, **request):
        mode = request["mode"]
        if mode == "time":
            return [({"mode": "time"}, None), (self.source, request)]
        kwargs = self._snap_kwargs
        time_data = self.source.get_data(mode="time", start=request.get("start"), stop=request.get("stop"))
        if time_data is None or not time_data.get("time"):
            return [({"empty": True, "mode": mode}, None)]
        start, stop = time_data["time"][0], time_data["time"][-1]
        if self.frequency is None:
            request["start"] = self.period[0]
            request["stop"] = stop
        else:
            request["start"] = _ts_to_dt(_get_bin_start(start, **kwargs), self.timezone)
            request["stop"] = stop + MICROSECOND  # bins are closed on the right
        kwargs.update(mode=mode, start=start, stop=stop)
        if mode == "vals":
            kwargs["dtype"] = np.dtype(self.dtype).str
            kwargs["statistic"] = self.statistic
        time_request = {"mode": "time", "start": request["start"], "stop": request["stop"]}
        return [(kwargs, None), (self.source, time_request), (self.source, request)]

    @staticmethod
    def process(process_kwargs, time_data=None, data=None):
        mode = process_kwargs["mode"]
        empty = None if mode == "vals" else {mode: []}
        if process_kwargs.get("empty"):
            return empty
        if mode == "time":
            return time_data
        if time_data is None or not time_data.get("time"):
            return empty
        frequency, timezone = process_kwargs["frequency"], process_kwargs["timezone"]
        if frequency is None:
            times = pd.Series(index=time_data["time"], dtype=float).tz_localize("UTC").tz_convert(timezone)
            indices = {None: range(len(times))}
        else:
            times, indices = _resample_indices(time_data["time"], frequency, process_kwargs["closed"],
                                               process_kwargs["label"], timezone)
        start_ts = _dt_to_ts(process_kwargs["start"], timezone)
        stop_ts = _dt_to_ts(process_kwargs["stop"], timezone)

        if mode == "meta":
            if data is None or "meta" not in data:
                return {"meta": []}
            meta, result = data["meta"], []
            for members in indices.values():
                for length in range(1, len(members) + 1):
                    upto = members[:length]
                    ts = times.index[upto[-1]]
                    if ts < start_ts or (stop_ts is not None and ts > stop_ts):
                        continue
                    result.append([meta[i] for i in upto])
            return {"meta": result}

        if data is None or "values" not in data:
            return None
        values = data["values"]
        if values.shape[0] != len(times):
            raise RuntimeError("Shape of raster does not match number of timestamps")
        statistic, _ = parse_percentile_statistic(process_kwargs["statistic"])
        dtype = np.dtype(process_kwargs["dtype"])
        wanted = np.asarray((times.index >= start_ts) & (times.index <= stop_ts))
        first = int(np.where(wanted)[0][0])
        groups = [list(members) for members in indices.values()]
        offsets, frames = _bins_to_arrays(groups)
        out_frame = np.where(wanted[frames], frames - first, -1).astype(np.int32)
        holder, nodata_ptr, has_nodata = _nodata_arg(values, data["no_data_value"])
        out = _run(
            values, (int(wanted.sum()), values.shape[1], values.shape[2]), dtype,
            lambda lib, src, dst, stream: lib.gm_temporal_cumulative(
                src, dst, nodata_ptr, has_nodata, _STAT_CODES[statistic], offsets.ctypes.data,
                frames.ctypes.data, out_frame.ctypes.data, len(groups), stream),
        )
        return {"values": out, "no_data_value": get_dtype_max(dtype)}
