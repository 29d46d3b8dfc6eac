"""Test-side blocks with the behaviour of the reference's test factories
(reference: dask_geomodeling/tests/factories.py:23-282), so that the reference's own unit
tests can be replayed through the CUDA blocks.

``MockRaster`` answers requests on the CPU (it is a *source* in those tests, not part of the
path under test); ``MockGeometry`` is the package's ``MemoryGeometrySource``.
"""
import math

import numpy as np
from scipy import ndimage

from dask_geomodeling_b200.geometry import MemoryGeometrySource as MockGeometry  # noqa: F401
from dask_geomodeling_b200.raster import RasterBlock
from dask_geomodeling_b200.utils import get_dtype_max


class MockRaster(RasterBlock):
    """Constant (uint8) or array-valued raster; for arrays the request bbox is read as
    indices into the array (row index = y, no north-up flip), out-of-range cells are no
    data, and a request of another shape is zoomed (factories.py:61-152)."""

    def __init__(self, origin=None, timedelta=None, bands=None, value=1, projection="EPSG:3857",
                 temporal=None):
        if temporal is None:
            temporal = timedelta is not None
        super().__init__(origin, timedelta, bands, value, projection, temporal)

    origin = property(lambda self: self.args[0])
    timedelta = property(lambda self: self.args[1])
    bands = property(lambda self: self.args[2])
    value = property(lambda self: self.args[3])
    projection = property(lambda self: self.args[4])
    temporal = property(lambda self: self.args[5])
    geometry = None

    @property
    def dtype(self):
        return getattr(self.value, "dtype", np.dtype("u1"))

    @property
    def fillvalue(self):
        return get_dtype_max(self.dtype)

    @property
    def period(self):
        if None in (self.origin, self.bands, self.timedelta):
            return None
        return self.origin, self.origin + (self.bands - 1) * self.timedelta

    @property
    def extent(self):
        if self.value is None:
            return None
        if np.isscalar(self.value):
            return 0, 0, 1, 1
        h, w = self.value.shape
        return 0, 0, w, h

    @property
    def geo_transform(self):
        x1, _, _, y2 = self.extent
        return x1, 1, 0, y2, 0, -1

    def get_sources_and_requests(self, **request):
        return [(self.args, None), (request, None)]

    @staticmethod
    def process(args, request):
        origin, delta, bands, value, _, _ = args
        if origin is None or delta is None or bands is None:
            return None
        step = delta.total_seconds()
        start, stop = request.get("start"), request.get("stop")
        if start is None:
            lo, hi = bands - 1, bands
        elif stop is None:
            lo = min(max(int(round((start - origin).total_seconds() / step)), 0), bands - 1)
            hi = lo + 1
        else:
            lo = max(int(math.ceil((start - origin).total_seconds() / step)), 0)
            hi = min(int(math.floor((stop - origin).total_seconds() / step)) + 1, bands)
        depth = hi - lo
        if depth <= 0:
            return None
        mode = request["mode"]
        if mode == "time":
            return {"time": [origin + i * delta for i in range(lo, hi)]}
        if mode == "meta":
            return {"meta": ["Testmeta for band {}".format(i) for i in range(lo, hi)]}
        if mode != "vals":
            raise ValueError('Invalid mode "{}"'.format(mode))
        height, width = request.get("height", 1), request.get("width", 1)
        if not hasattr(value, "shape"):
            return {"values": np.full((depth, height, width), value, dtype="u1"), "no_data_value": 255}

        fill = get_dtype_max(value.dtype)
        x1, y1, x2, y2 = (int(round(v)) for v in request.get("bbox", (0, 0, width, height)))
        if x1 == x2 or y1 == y2:
            inside = 0 <= x1 < value.shape[1] and 0 <= y1 < value.shape[0]
            frame = value[y1:y1 + 1, x1:x1 + 1] if inside else np.array([[255]], dtype="u1")
        else:
            cx1, cy1 = max(x1, 0), max(y1, 0)
            cx2, cy2 = min(x2, value.shape[1]), min(y2, value.shape[0])
            frame = np.pad(value[cy1:cy2, cx1:cx2], ((cy1 - y1, y2 - cy2), (cx1 - x1, x2 - cx2)),
                           mode="constant", constant_values=fill)
            if frame.shape != (height, width):
                zoom = (height / frame.shape[0], width / frame.shape[1])
                missing = ndimage.zoom((frame == fill).astype(float), zoom) > 0.5
                frame[frame == fill] = 0
                frame = ndimage.zoom(frame, zoom)
                frame[missing] = fill
        result = np.repeat(frame[np.newaxis], depth, axis=0)
        result[~np.isfinite(result)] = fill
        return {"values": result, "no_data_value": fill}
