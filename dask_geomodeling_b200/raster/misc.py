"""Clip, Mask, MaskBelow, Step, Classify, Reclassify, Rasterize(WKT).

Drop-in for the reference's raster/misc.py: identical constructors,
attributes and ``process`` signatures; the per-pixel work runs in the CUDA
evaluator (``_lowering``), polygon burning in the CUDA scanline rasteriser
(``utils.rasterize_geoseries``).
"""
import numpy as np

from .. import _native

from .. import utils
from .._compat import config
from . import _lowering
from .base import BaseSingle, RasterBlock

__all__ = ["Clip", "Classify", "Reclassify", "Mask", "MaskBelow", "Step", "Rasterize", "RasterizeWKT"]


def _number(x):
    if not isinstance(x, (float, int)):
        raise TypeError("'{}' object is not allowed".format(type(x)))
    return x


class Clip(BaseSingle):
    """Keep ``store`` where ``source`` has data (or is True); elsewhere no data
    (reference: raster/misc.py:30-166)."""

    def __init__(self, store, source):
        if not isinstance(source, RasterBlock):
            raise TypeError("'{}' object is not allowed".format(type(store)))
        if store.temporal and not source.temporal:
            raise ValueError(
                "The values raster is temporal while the clipping mask is not. Consider using Snap."
            )
        if source.temporal and not store.temporal:
            raise ValueError(
                "The clipping mask is temporal while the values raster is not. Consider using Snap."
            )
        if store.temporal and store.timedelta != source.timedelta:
            raise ValueError(
                "Time resolution of the clipping mask does not match that of "
                "the values raster. Consider using Snap."
            )
        super(Clip, self).__init__(store, source)

    @property
    def source(self):
        return self.args[1]

    def get_sources_and_requests(self, **request):
        period = self.period
        nothing = [(None, None), (None, None)]
        if period is None:
            return nothing
        first, last = period
        start = request.get("start")
        stop = request.get("stop")
        if start is None:
            start = last
        if stop is not None:
            if stop < first or start > last:
                return nothing
            request["stop"] = min(max(stop, first), last)
        request["start"] = min(max(start, first), last)
        return [(arg, request) for arg in self.args]

    process = staticmethod(_lowering.clip_process)

    @property
    def extent(self):
        boxes = [a.extent for a in self.args]
        if any(b is None for b in boxes):
            return None
        x1, y1 = max(b[0] for b in boxes), max(b[1] for b in boxes)
        x2, y2 = min(b[2] for b in boxes), min(b[3] for b in boxes)
        return None if (x2 <= x1 or y2 <= y1) else (x1, y1, x2, y2)

    @property
    def geometry(self):
        a, b = (x.geometry for x in self.args)
        if a is None or b is None:
            return None
        common = utils.Extent.from_geometry(a).intersection(utils.Extent.from_geometry(b))
        return None if common is None else common.as_geometry()

    @property
    def period(self):
        periods = [a.period for a in self.args]
        if any(p is None for p in periods):
            return None
        start, stop = max(p[0] for p in periods), min(p[1] for p in periods)
        return None if stop < start else (start, stop)


class Mask(BaseSingle):
    """Replace every data cell by ``value``; no data cells get 0 (1 if value is 0)
    (reference: raster/misc.py:169-222)."""

    def __init__(self, store, value):
        super(Mask, self).__init__(store, _number(value))

    @property
    def value(self):
        return self.args[1]

    @property
    def fillvalue(self):
        return 1 if self.value == 0 else 0

    @staticmethod
    def _dtype_from_value(value):
        if isinstance(value, float):
            return np.dtype("float32")
        return utils.get_uint_dtype(value) if value >= 0 else utils.get_int_dtype(value)

    @property
    def dtype(self):
        return self._dtype_from_value(self.value)

    process = staticmethod(_lowering.mask_process)


class MaskBelow(BaseSingle):
    """Cells below ``value`` become no data (reference: raster/misc.py:225-251)."""

    def __init__(self, store, value):
        super(MaskBelow, self).__init__(store, _number(value))

    process = staticmethod(_lowering.maskbelow_process)


class Step(BaseSingle):
    """``left`` below ``value``, ``at`` on it, ``right`` above it
    (reference: raster/misc.py:254-328)."""

    def __init__(self, store, left=0, right=1, value=0, at=None):
        if at is None:
            at = (left + right) / 2
        for x in (left, right, value, at):
            _number(x)
        super(Step, self).__init__(store, left, right, value, at)

    left = property(lambda self: self.args[1])
    right = property(lambda self: self.args[2])
    value = property(lambda self: self.args[3])
    at = property(lambda self: self.args[4])

    process = staticmethod(_lowering.step_process)


class Classify(BaseSingle):
    """Index of the bin (``np.digitize``) every cell falls in
    (reference: raster/misc.py:331-399)."""

    def __init__(self, store, bins, right=False):
        if not isinstance(store, RasterBlock):
            raise TypeError("'{}' object is not allowed".format(type(store)))
        if not hasattr(bins, "__iter__"):
            raise TypeError("'{}' object is not allowed".format(type(bins)))
        edges = np.asarray(bins)
        if edges.ndim != 1:
            raise TypeError("'bins' should be one-dimensional")
        if not np.issubdtype(edges.dtype, np.number):
            raise TypeError("'bins' should be numeric")
        steps = np.diff(bins)
        if not np.all(steps > 0) or np.all(steps < 0):
            raise TypeError("'bins' should be monotonic")
        super(Classify, self).__init__(store, edges.tolist(), right)

    bins = property(lambda self: self.args[1])
    right = property(lambda self: self.args[2])

    @property
    def dtype(self):
        # n edges give n + 1 classes, plus one value for 'no data'
        return utils.get_uint_dtype(len(self.bins) + 2)

    @property
    def fillvalue(self):
        return utils.get_dtype_max(self.dtype)

    process = staticmethod(_lowering.classify_process)


class Reclassify(BaseSingle):
    """Map integer cell values through a list of ``[from, to]`` pairs
    (reference: raster/misc.py:402-515)."""

    def __init__(self, store, data, select=False):
        dtype = store.dtype
        if dtype != bool and not np.issubdtype(dtype, np.integer):
            raise TypeError("The store must be of boolean or integer datatype")
        if not hasattr(data, "__iter__"):
            raise TypeError("'{}' object is not allowed".format(type(data)))
        try:
            source, target = self._data_as_ndarray(data)
        except ValueError:
            raise ValueError("Please supply a list of [from, to] values")
        if source.dtype != bool and not np.issubdtype(source.dtype, np.integer):
            raise TypeError("Cannot reclassify from value with type '{}'".format(source.dtype))
        if len(np.unique(source)) != len(source):
            raise ValueError("There are duplicates in the reclassify values")
        if not np.issubdtype(target.dtype, np.number):
            raise TypeError("Cannot reclassify to value with type '{}'".format(target.dtype))
        pairs = [list(pair) for pair in zip(source.tolist(), target.tolist())]
        if select is not True and select is not False:
            raise TypeError("'{}' object is not allowed".format(type(select)))
        super().__init__(store, pairs, select)

    @staticmethod
    def _data_as_ndarray(data):
        source, target = zip(*data)
        return np.asarray(source), np.asarray(target)

    data = property(lambda self: self.args[1])
    select = property(lambda self: self.args[2])

    @property
    def dtype(self):
        return self._data_as_ndarray(self.data)[1].dtype

    @property
    def fillvalue(self):
        return utils.get_dtype_max(self.dtype)

    def get_sources_and_requests(self, **request):
        process_kwargs = {
            "dtype": self.dtype.str,
            "fillvalue": self.fillvalue,
            "data": self.data,
            "select": self.select,
        }
        return [(self.store, request), (process_kwargs, None)]

    process = staticmethod(_lowering.reclassify_process)


class _NonTemporalRaster(RasterBlock):
    """Attributes shared by rasters that are generated from geometries."""

    extent = None
    timedelta = None
    temporal = False
    geometry = None
    geo_transform = None

    @property
    def period(self):
        return (self.DEFAULT_ORIGIN,) * 2


class Rasterize(_NonTemporalRaster):
    """Burn a column of a GeometryBlock into a raster
    (reference: raster/misc.py:518-709)."""

    projection = None

    def __init__(self, source, column_name=None, dtype=None, limit=None):
        from ..geometry.base import GeometryBlock

        if not isinstance(source, GeometryBlock):
            raise TypeError("'{}' object is not allowed".format(type(source)))
        if column_name is not None and not isinstance(column_name, str):
            raise TypeError("'{}' object is not allowed".format(type(column_name)))
        if dtype is None:
            dtype = "bool" if column_name is None else "int32"
        else:
            dtype = str(np.dtype(dtype))
        if limit and not isinstance(limit, int):
            raise TypeError("'{}' object is not allowed".format(type(limit)))
        if limit and limit < 1:
            raise ValueError("Limit should be greater than 1")
        super(Rasterize, self).__init__(source, column_name, dtype, limit)

    source = property(lambda self: self.args[0])
    column_name = property(lambda self: self.args[1])
    limit = property(lambda self: self.args[3])

    @property
    def dtype(self):
        return np.dtype(self.args[2])

    @property
    def fillvalue(self):
        return None if self.dtype == bool else utils.get_dtype_max(self.dtype)

    def get_sources_and_requests(self, **request):
        mode = request["mode"]
        if mode == "time":
            return [(self.period[-1], None), ({"mode": "time"}, None)]
        if mode == "meta":
            return [(None, None), ({"mode": "meta"}, None)]
        if mode != "vals":
            raise ValueError("Unknown mode '{}'".format(mode))

        x1, y1, x2, y2 = request["bbox"]
        width, height = request["width"], request["height"]
        if x2 == x1 and y2 == y1:
            min_size = None  # point request
        elif x1 < x2 and y1 < y2:
            min_size = min((x2 - x1) / width, (y2 - y1) / height)
        else:
            raise ValueError("Invalid bbox ({})".format(request["bbox"]))
        limit = self.limit
        if limit is None:
            limit = config.get("geomodeling.geometry-limit")
        geom_request = {
            "mode": "intersects",
            "geometry": utils.box(*request["bbox"]),
            "projection": request["projection"],
            "min_size": min_size,
            "limit": limit,
            "start": request.get("start"),
            "stop": request.get("stop"),
        }
        process_kwargs = {
            "mode": "vals",
            "column_name": self.column_name,
            "dtype": self.dtype,
            "no_data_value": self.fillvalue,
            "width": width,
            "height": height,
            "bbox": request["bbox"],
        }
        return [(self.source, geom_request), (process_kwargs, None)]

    @staticmethod
    def process(data, process_kwargs):
        mode = process_kwargs["mode"]
        if mode == "time":
            return {"time": [data]}
        if mode == "meta":
            return {"meta": [None]}
        column_name = process_kwargs["column_name"]
        height, width = process_kwargs["height"], process_kwargs["width"]
        no_data_value = process_kwargs["no_data_value"]
        dtype = process_kwargs["dtype"]
        features = data["features"]

        if column_name is None:
            values = None
        elif column_name in features:
            values = features[column_name]
        elif features.index.name == column_name:
            values = features.index.to_series()
        else:
            values = False
        if len(features) == 0 or values is False:
            empty = np.full((1, height, width), no_data_value, dtype=dtype)
            return {"values": empty, "no_data_value": no_data_value}

        # the geometry source's CSR form of these geometries, if the frame still carries it
        from ..geometry.sources import prepared_soup_of

        soup = None
        prepared = prepared_soup_of(features)
        if prepared is not None:
            full, positions, _ = prepared
            soup = full if len(positions) == full.n_polygons else full.subset(positions)
        burned = utils.rasterize_geoseries(
            geoseries=features["geometry"] if "geometry" in features else None,
            values=values,
            bbox=process_kwargs["bbox"],
            projection=data["projection"],
            height=height,
            width=width,
            soup=soup,
        )
        raw = burned["values"]
        if _native.is_device(raw):
            if raw.dtype == dtype and burned["no_data_value"] == no_data_value:
                return {"values": raw, "no_data_value": no_data_value}    # stays in HBM
            raw = np.asarray(raw)
        with np.errstate(over="ignore", under="ignore"):
            result = raw.astype(dtype)
        if burned["no_data_value"] != no_data_value:
            result[raw == burned["no_data_value"]] = no_data_value
        return {"values": result, "no_data_value": no_data_value}


class RasterizeWKT(_NonTemporalRaster):
    """Boolean raster of one WKT geometry (reference: raster/misc.py:712-830)."""

    dtype = np.dtype("bool")
    fillvalue = None

    def __init__(self, wkt, projection):
        if not isinstance(wkt, str):
            raise TypeError("'{}' object is not allowed".format(type(wkt)))
        if not isinstance(projection, str):
            raise TypeError("'{}' object is not allowed".format(type(projection)))
        try:
            utils.shapely_from_wkt(wkt)
        except utils.WKTReadingError:
            raise ValueError("The provided geometry is not a valid WKT")
        super().__init__(wkt, projection)

    wkt = property(lambda self: self.args[0])
    projection = property(lambda self: self.args[1])

    @property
    def extent(self):
        geometry = utils.shapely_from_wkt(self.wkt)
        return tuple(utils.Extent(geometry.bounds, self.projection).transformed("EPSG:4326").bbox)

    @property
    def geometry(self):
        return utils.shapely_from_wkt(self.wkt)

    def get_sources_and_requests(self, **request):
        mode = request["mode"]
        if mode == "time":
            data = self.period[-1]
        elif mode == "meta":
            data = None
        elif mode == "vals":
            data = {"wkt": self.wkt, "projection": self.projection}
        else:
            raise ValueError("Unknown mode '{}'".format(mode))
        return [(data, None), (request, None)]

    @staticmethod
    def process(data, request):
        mode = request["mode"]
        if mode == "time":
            return {"time": [data]}
        if mode == "meta":
            return {"meta": [None]}
        geometry = utils.shapely_from_wkt(data["wkt"])
        if not utils.same_projection(data["projection"], request["projection"]):
            geometry = utils.shapely_transform(geometry, data["projection"], request["projection"])
        x1, y1, x2, y2 = request["bbox"]
        gx1, gy1, gx2, gy2 = geometry.bounds if not geometry.is_empty else (1, 1, 0, 0)
        if geometry.is_empty or gx2 < x1 or gx1 > x2 or gy2 < y1 or gy1 > y2:
            empty = np.full((1, request["height"], request["width"]), False, dtype=bool)
            return {"values": empty, "no_data_value": None}
        return utils.rasterize_geoseries(
            geoseries=[geometry],
            bbox=request["bbox"],
            projection=request["projection"],
            height=request["height"],
            width=request["width"],
        )
