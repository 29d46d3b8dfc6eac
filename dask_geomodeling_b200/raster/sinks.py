"""Raster sinks: tiles of a view written as GeoTIFF files, merged into a VRT
(reference: raster/sinks.py:18-204).

The tile values come out of the CUDA path (``RasterTiler`` cuts the request, every tile is
evaluated on the device and downloaded once); encoding them is host work
(``dask_geomodeling_b200/geotiff.py``: 256 x 256 deflate tiles on a thread pool, no GDAL).
"""
import glob
import os

import numpy as np

from .. import _native, geotiff, utils
from .._compat import tokenize
from .base import BaseSingle, RasterBlock
from .parallelize import RasterTiler

__all__ = ["RasterFileSink", "to_file"]


class RasterFileSink(BaseSingle):
    """Write raster data to GeoTIFF files in a specified directory.

    Use RasterFileSink.merge_files to merge tiles into a VRT file.

    Args:
      source (RasterBlock): The raster block the data is coming from.
      url (str): The target directory to put the files in. If relative, it is taken relative to
        the geomodeling.root setting.
    """

    def __init__(self, source, url):
        if not isinstance(source, RasterBlock):
            raise TypeError("'{}' object is not allowed".format(type(source)))
        super().__init__(source, utils.safe_file_url(url))

    @property
    def url(self):
        return self.args[1]

    def get_sources_and_requests(self, **request):
        if request["mode"] != "vals":
            return [(self.store, request), ({}, None)]
        process_kwargs = {
            "url": self.url,
            "hash": tokenize(request)[:7],
            "bbox": request["bbox"],
            "projection": request["projection"],
        }
        return [(self.store, request), (process_kwargs, None)]

    @staticmethod
    def process(data, process_kwargs):
        if not process_kwargs:
            return data    # non-vals mode: forward data as-is
        if data is None or "values" not in data:
            return None
        values = data["values"]
        if _native.is_device(values):
            values = np.asarray(values)
        no_data_value = data["no_data_value"]
        if values.ndim != 3 or values.shape[0] != 1:
            raise ValueError("Expected a single-band raster (shape (1, H, W)), got shape {}".format(values.shape))
        band_data = values[0]
        if no_data_value is not None and np.all(band_data == no_data_value):
            return None    # nothing but no data: no file
        height, width = band_data.shape
        path = utils.safe_abspath(process_kwargs["url"])
        os.makedirs(path, exist_ok=True)
        x1, y1, x2, y2 = process_kwargs["bbox"]
        geo_transform = (x1, (x2 - x1) / width, 0, y2, 0, -(y2 - y1) / height)
        geotiff.write_geotiff(os.path.join(path, process_kwargs["hash"] + ".tif"), band_data, geo_transform,
                              process_kwargs["projection"], no_data_value)
        return None

    @staticmethod
    def merge_files(path, target):
        """Merge GeoTIFF files (the output of this Block) into a VRT file.

        Args:
          path (str): The source directory containing .tif files.
          target (str): The target .vrt file path.
        """
        path = utils.safe_abspath(path)
        target = utils.safe_abspath(target)
        if os.path.exists(target):
            raise IOError("Target '{}' already exists".format(target))
        source_paths = glob.glob(os.path.join(path, "*.tif"))
        if len(source_paths) == 0:
            raise IOError("No source .tif files found in '{}'".format(path))
        geotiff.write_vrt(target, source_paths)


def to_file(source, url, tile_size, **request):
    """Export data from a RasterBlock to disk: tiled GeoTIFFs merged into a VRT at ``url``
    (reference raster/sinks.py:148-204; same defaults for projection, bbox, width and height)."""
    request["mode"] = "vals"
    if "projection" not in request:
        if source.projection is None:
            raise ValueError("Cannot determine the projection from the source raster. "
                             "Please provide a 'projection' argument.")
        request["projection"] = source.projection
    if "bbox" not in request:
        if source.geometry is None:
            raise ValueError("Cannot determine the extent from the source raster. "
                             "Please provide a 'bbox' argument.")
        if hasattr(source.geometry, "GetEnvelope"):
            x1, x2, y1, y2 = source.geometry.GetEnvelope()
        else:
            x1, y1, x2, y2 = source.geometry.bounds
        request["bbox"] = x1, y1, x2, y2
    if "width" not in request or "height" not in request:
        if source.geo_transform is None:
            raise ValueError("Cannot determine the pixel size from the source raster. "
                             "Please provide 'width' and 'height' arguments.")
        geo_transform = source.geo_transform
        x1, y1, x2, y2 = request["bbox"]
        request["width"] = int(round((x2 - x1) / abs(float(geo_transform[1]))))
        request["height"] = int(round((y2 - y1) / abs(float(geo_transform[5]))))
    path = utils.safe_abspath(url)
    if os.path.isdir(path):
        path = os.path.join(path, "output.vrt")
    tiles_dir = os.path.join(os.path.split(path)[0], "tiles")
    sink = RasterFileSink(source, tiles_dir)
    RasterTiler(sink, tile_size).get_data(**request)
    RasterFileSink.merge_files(tiles_dir, path)
