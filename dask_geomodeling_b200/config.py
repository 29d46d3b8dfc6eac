"""Configuration defaults in the ``geomodeling`` namespace of dask.config
(same keys as the reference's config.py:4-12, plus the device knobs)."""
import os

from ._compat import config

defaults = {
    "root": os.getcwd(),
    "strict-file-paths": False,
    "raster-limit": 12 * (1024 ** 2),
    "raster-limit-timesteps": 65536,
    "geometry-limit": 10000,
    # additions of the CUDA build
    "fuse": True,             # fuse element-wise sub-graphs into single launches
    "device-resident": True,  # keep rasters in HBM between the tasks of one compute()
    "pin-sources": True,      # page-lock MemorySource arrays for full-speed uploads
}

config.update_defaults({"geomodeling": defaults})
