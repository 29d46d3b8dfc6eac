"""RasterBlock contract (drop-in for the reference's raster/base.py).

Requests: ``mode`` ('vals' | 'time' | 'meta'), ``bbox``, ``projection``,
``width``, ``height``, ``start``, ``stop``.  Responses: None, or a dict with
``values`` (bands, height, width) + ``no_data_value``, or ``time`` / ``meta``
lists (reference: raster/base.py:27-46).  Every RasterBlock exposes
``period, timedelta, extent, dtype, fillvalue, geometry, projection,
geo_transform, temporal`` (raster/base.py:15-23).
"""
from datetime import datetime as Datetime

from ..core import Block

__all__ = ["RasterBlock", "BaseSingle"]


def _operator(block_name, reflected=False):
    def method(self, other):
        from . import elemwise

        klass = getattr(elemwise, block_name)
        return klass(other, self) if reflected else klass(self, other)

    method.__name__ = "__{}__".format(block_name.lower())
    return method


class RasterBlock(Block):
    """Base of all (temporal) raster views; python operators build blocks
    (raster/base.py:96-174)."""

    DEFAULT_ORIGIN = Datetime(1970, 1, 1, 0, 0)

    __add__ = _operator("Add")
    __sub__ = _operator("Subtract")
    __mul__ = _operator("Multiply")
    __truediv__ = _operator("Divide")
    __pow__ = _operator("Power")
    __eq__ = _operator("Equal")
    __ne__ = _operator("NotEqual")
    __gt__ = _operator("Greater")
    __ge__ = _operator("GreaterEqual")
    __lt__ = _operator("Less")
    __le__ = _operator("LessEqual")
    __and__ = _operator("And")
    __or__ = _operator("Or")
    __xor__ = _operator("Xor")
    __hash__ = Block.__hash__  # defining __eq__ would otherwise drop hashing

    def __neg__(self):
        from .elemwise import Multiply

        return Multiply(self, -1)

    def __invert__(self):
        from .elemwise import Invert

        return Invert(self)

    def to_file(self, *args, **kwargs):
        """Export this block to tiled GeoTIFFs + a VRT: ``to_file(url, tile_size, **request)``
        (raster/base.py:51-73, raster/sinks.py:148-204)."""
        from .sinks import to_file

        return to_file(self, *args, **kwargs)

    def __len__(self):
        """Number of frames on the time axis."""
        period = self.period
        if period is None:
            return 0
        start, stop = period
        if start == stop:
            return 1
        delta = self.timedelta
        if delta is None:  # non-equidistant: ask the source
            return len(self.get_data(mode="time", start=start, stop=stop)["time"])
        return int((stop - start).total_seconds() / delta.total_seconds()) + 1


class BaseSingle(RasterBlock):
    """A block that transforms one raster (``store``) and inherits its attributes."""

    def __init__(self, store, *args):
        if not isinstance(store, RasterBlock):
            raise TypeError("'{}' object is not allowed".format(type(store)))
        super(BaseSingle, self).__init__(store, *args)

    @property
    def store(self):
        return self.args[0]

    def __len__(self):
        return len(self.store)


def _delegate(attribute):
    return property(lambda self: getattr(self.store, attribute))


for _name in ("extent", "period", "timedelta", "temporal", "dtype", "fillvalue", "geometry",
              "projection", "geo_transform"):
    setattr(BaseSingle, _name, _delegate(_name))
