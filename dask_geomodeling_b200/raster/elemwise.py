"""Element-wise raster blocks, evaluated by the CUDA bytecode evaluator.

Same classes, constructor checks, dtype/fillvalue rules and ``process``
call signatures as the reference's raster/elemwise.py; the NumPy body of each
``process`` (raster/elemwise.py:235-299, :570-575, :601-638, :726-757) is
replaced by a one-node program for ``gm_eval_program``.  When a view is
evaluated through ``get_data`` whole chains of these blocks are fused into a
single launch (core/fusion.py).
"""
import numpy as np

from ..utils import Extent, GeoTransform, get_dtype_max
from . import _lowering
from .base import BaseSingle, RasterBlock

__all__ = [
    "Add", "Subtract", "Multiply", "Divide", "Power", "FillNoData", "Equal", "NotEqual",
    "Greater", "GreaterEqual", "Less", "LessEqual", "Invert", "And", "Or", "Xor", "IsData",
    "IsNoData", "Exp", "Log", "Log10",
]


def _intersect_boxes(boxes):
    """Common part of (x1, y1, x2, y2) boxes; None if any is None or they are disjoint."""
    if any(b is None for b in boxes):
        return None
    x1, y1 = max(b[0] for b in boxes), max(b[1] for b in boxes)
    x2, y2 = min(b[2] for b in boxes), min(b[3] for b in boxes)
    return None if (x2 <= x1 or y2 <= y1) else (x1, y1, x2, y2)


class BaseElementwise(RasterBlock):
    """Pixel-wise combination of aligned rasters and constants.  Spatial and
    temporal attributes are the intersection of those of the raster arguments."""

    def __init__(self, *args):
        super(BaseElementwise, self).__init__(*args)
        rasters = self._sources
        if len(rasters) > 1:
            first = rasters[0]
            if any(r.temporal != first.temporal for r in rasters[1:]):
                raise ValueError("Temporal properties of input rasters do not match.")
            if first.temporal and first.timedelta is not None:
                if any(r.timedelta not in (None, first.timedelta) for r in rasters[1:]):
                    raise ValueError("Time resolutions of input rasters are not equal.")

    @property
    def _sources(self):
        return [a for a in self.args if isinstance(a, RasterBlock)]

    def get_sources_and_requests(self, **request):
        start, stop = request.get("start"), request.get("stop")
        if start is not None and stop is not None:
            period = self.period
            if period is not None:  # clamp so that all sources return aligned frames
                request["start"] = max(start, period[0])
                request["stop"] = min(stop, period[1])
        process_kwargs = {"dtype": self.dtype.name, "fillvalue": self.fillvalue}
        return [(process_kwargs, None)] + [(arg, request) for arg in self.args]

    @property
    def dtype(self):
        dtype = np.result_type(*self.args)
        if dtype == bool or np.issubdtype(dtype, np.integer):
            return np.result_type(dtype, np.int32)
        if np.issubdtype(dtype, np.floating):
            return np.result_type(dtype, np.float32)
        return dtype

    @property
    def fillvalue(self):
        dtype = self.dtype
        return None if dtype == bool else get_dtype_max(dtype)

    @property
    def temporal(self):
        return self._sources[0].temporal

    @property
    def timedelta(self):
        deltas = [r.timedelta for r in self._sources]
        return None if any(d is None for d in deltas) else deltas[0]

    @property
    def period(self):
        periods = [r.period for r in self._sources]
        if any(p is None for p in periods):
            return None
        start, stop = max(p[0] for p in periods), min(p[1] for p in periods)
        return None if stop < start else (start, stop)

    @property
    def extent(self):
        return _intersect_boxes([r.extent for r in self._sources])

    @property
    def geometry(self):
        geometries = [r.geometry for r in self._sources]
        if any(g is None for g in geometries):
            return None
        if len(geometries) == 1:
            return geometries[0]
        extent = Extent.from_geometry(geometries[0])
        for g in geometries[1:]:
            extent = extent.intersection(Extent.from_geometry(g))
            if extent is None:
                return None
        return extent.as_geometry()

    @property
    def projection(self):
        projections = {r.projection for r in self._sources}
        return projections.pop() if len(projections) == 1 else None

    @property
    def geo_transform(self):
        first = self._sources[0].geo_transform
        if first is None:
            return None
        first = GeoTransform(first)
        for r in self._sources[1:]:
            other = r.geo_transform
            if other is None or not first.aligns_with(other):
                return None
        return first


class BaseMath(BaseElementwise):
    def __init__(self, a, b):
        for x in (a, b):
            if not isinstance(x, (RasterBlock, np.ndarray, float, int)):
                raise TypeError("'{}' object is not allowed".format(type(x)))
        super(BaseMath, self).__init__(a, b)


class BaseComparison(BaseMath):
    @property
    def dtype(self):
        return np.dtype("bool")


class BaseLogic(BaseElementwise):
    def __init__(self, a, b):
        for x in (a, b):
            if isinstance(x, (RasterBlock, np.ndarray)):
                if x.dtype != np.dtype("bool"):
                    raise TypeError("inputs must have boolean dtypes")
            elif not isinstance(x, bool):
                raise TypeError("'{}' object is not allowed".format(type(x)))
        super(BaseLogic, self).__init__(a, b)

    @property
    def dtype(self):
        return np.dtype("bool")

    @property
    def fillvalue(self):
        return None


class Add(BaseMath):
    """a + b; no data where either input has no data."""

    process = staticmethod(_lowering.math_process("add"))


class Subtract(BaseMath):
    """a - b."""

    process = staticmethod(_lowering.math_process("subtract"))


class Multiply(BaseMath):
    """a * b."""

    process = staticmethod(_lowering.math_process("multiply"))


class Divide(BaseMath):
    """a / b in floating point (at least float32); x/0 becomes no data."""

    process = staticmethod(_lowering.math_process("divide"))

    @property
    def dtype(self):
        return np.result_type(np.float32, *self.args)


class Power(BaseMath):
    """a ** b; a negative integer exponent is taken as a float."""

    process = staticmethod(_lowering.math_process("power"))

    def __init__(self, a, b):
        if isinstance(b, int) and b < 0:
            b = float(b)
        super(Power, self).__init__(a, b)


class Equal(BaseComparison):
    """a == b; False where either input has no data."""

    process = staticmethod(_lowering.math_process("equal"))


class NotEqual(BaseComparison):
    """a != b; True where either input has no data."""

    process = staticmethod(_lowering.math_process("not_equal"))


class Greater(BaseComparison):
    """a > b; False where either input has no data."""

    process = staticmethod(_lowering.math_process("greater"))


class GreaterEqual(BaseComparison):
    """a >= b; False where either input has no data."""

    process = staticmethod(_lowering.math_process("greater_equal"))


class Less(BaseComparison):
    """a < b; False where either input has no data."""

    process = staticmethod(_lowering.math_process("less"))


class LessEqual(BaseComparison):
    """a <= b; False where either input has no data."""

    process = staticmethod(_lowering.math_process("less_equal"))


class And(BaseLogic):
    process = staticmethod(_lowering.math_process("logical_and"))


class Or(BaseLogic):
    process = staticmethod(_lowering.math_process("logical_or"))


class Xor(BaseLogic):
    process = staticmethod(_lowering.math_process("logical_xor"))


class Invert(BaseSingle):
    """Swap True and False of a boolean raster."""

    def __init__(self, x):
        super(Invert, self).__init__(x)
        if x.dtype != np.dtype("bool"):
            raise TypeError("input block must have boolean dtype")

    process = staticmethod(_lowering.invert_process)

    @property
    def dtype(self):
        return np.dtype("bool")


class IsData(BaseSingle):
    """True where the raster has data."""

    def __init__(self, store):
        if store.dtype == np.dtype("bool"):
            raise TypeError("input block must not have boolean dtype")
        super(IsData, self).__init__(store)

    process = staticmethod(_lowering.isdata_process)

    @property
    def dtype(self):
        return np.dtype("bool")

    @property
    def fillvalue(self):
        return None


class IsNoData(IsData):
    """True where the raster has no data."""

    process = staticmethod(_lowering.isnodata_process)


class FillNoData(BaseElementwise):
    """Overlay rasters left to right; no data cells are transparent."""

    def __init__(self, *args):
        for arg in args:
            if not isinstance(arg, RasterBlock):
                raise TypeError("'{}' object is not allowed".format(type(arg)))
        super(FillNoData, self).__init__(*args)

    process = staticmethod(_lowering.fillnodata_process)


class BaseLogExp(BaseSingle):
    def __init__(self, x):
        if x.dtype == np.dtype("bool"):
            raise TypeError("input block must not have boolean dtype")
        super(BaseLogExp, self).__init__(x)

    def get_sources_and_requests(self, **request):
        process_kwargs = {"dtype": self.dtype.name, "fillvalue": self.fillvalue}
        return [(process_kwargs, None), (self.args[0], request)]

    @property
    def dtype(self):
        return np.result_type(np.float32, *self.args)

    @property
    def fillvalue(self):
        return get_dtype_max(self.dtype)


class Exp(BaseLogExp):
    """e ** x; overflow becomes no data."""

    process = staticmethod(_lowering.math_process("exp"))


class Log(BaseLogExp):
    """Natural logarithm; x <= 0 becomes no data."""

    process = staticmethod(_lowering.math_process("log"))


class Log10(BaseLogExp):
    """Base 10 logarithm; x <= 0 becomes no data."""

    process = staticmethod(_lowering.math_process("log10"))
