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
    # keep whole MemorySource arrays in HBM after their first request, up to this many bytes
    # in total (0 = off): later requests read the resident copy instead of uploading their
    # window again.  The source array must not be modified afterwards.
    "device-cache-bytes": 0,
}

config.update_defaults({"geomodeling": defaults})
