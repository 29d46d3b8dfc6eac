"""Raster blocks of the CUDA path: element-wise and misc blocks (fused into single launches),
stencils, temporal aggregation, reductions over several rasters (Max, Group), in-memory and file
sources, file sinks and the request tiler."""
from . import combine, elemwise, misc, parallelize, reduction, sinks, sources, spatial, temporal
from .base import RasterBlock

__all__ = ["RasterBlock"]
for _module in (elemwise, misc, sources, spatial, temporal, parallelize, reduction, combine, sinks):
    for _name in _module.__all__:
        globals()[_name] = getattr(_module, _name)
    __all__ += list(_module.__all__)
del _module, _name
