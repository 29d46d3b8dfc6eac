"""B200-native implementation of dask-geomodeling's per-tile raster compute path."""
from . import config  # NOQA
from .core import Block, construct  # NOQA
from . import raster  # NOQA
from . import geometry  # NOQA

__version__ = "0.1.0"
