"""Raster sinks: tiles of a view written as GeoTIFF files and merged into a VRT
(interface of the reference's raster/sinks.py:18-204: ``RasterFileSink(source, url)``,
``RasterFileSink.merge_files(path, target)``, ``to_file(source, url, tile_size, **request)``).

The tile values come out of the CUDA path -- ``RasterTiler`` cuts the request, every tile is
evaluated on the device and downloaded once -- and are encoded on the host by
``dask_geomodeling_b200/geotiff.py`` (256 x 256 deflate tiles on a thread pool, no GDAL).
"""
import glob
import os

import numpy as np

from .. import _native, geotiff, utils
from .._compat import tokenize
from .base import BaseSingle, RasterBlock
from .parallelize import RasterTiler

__all__ = ["RasterFileSink", "to_file"]


def _tile_path(directory_url, name):
    directory = utils.safe_abspath(directory_url)
    os.makedirs(directory, exist_ok=True)
    return os.path.join(directory, name + ".tif")


class RasterFileSink(BaseSingle):
    """Write the 'vals' response of ``source`` as one GeoTIFF per request into the directory
    ``url`` (relative paths start at the ``geomodeling.root`` setting).  The file is named after
    the request's token, so the tiles of a ``RasterTiler`` land next to each other; requests in
    other modes pass through.  ``merge_files`` turns such a directory into a VRT."""

    def __init__(self, source, url):
        if not isinstance(source, RasterBlock):
            raise TypeError("'{}' object is not allowed".format(type(source)))
        super().__init__(source, utils.safe_file_url(url))

    url = property(lambda self: self.args[1])

    def get_sources_and_requests(self, **request):
        target = {}
        if request["mode"] == "vals":
            target = {"url": self.url, "hash": tokenize(request)[:7],
                      "bbox": request["bbox"], "projection": request["projection"]}
        return [(self.store, request), (target, None)]

    @staticmethod
    def process(data, process_kwargs):
        if not process_kwargs:            # time / meta requests: nothing to write
            return data
        if data is None or "values" not in data:
            return None
        values, no_data_value = data["values"], data["no_data_value"]
        if _native.is_device(values):
            values = np.asarray(values)
        if values.ndim != 3 or values.shape[0] != 1:
            raise ValueError("Expected a single-band raster (shape (1, H, W)), got shape {}".format(values.shape))
        band = values[0]
        if no_data_value is not None and (band == no_data_value).all():
            return None                   # an empty tile leaves no file behind
        rows, cols = band.shape
        west, south, east, north = process_kwargs["bbox"]
        geo_transform = (west, (east - west) / cols, 0, north, 0, (south - north) / rows)
        geotiff.write_geotiff(_tile_path(process_kwargs["url"], process_kwargs["hash"]), band, geo_transform,
                              process_kwargs["projection"], no_data_value)
        return None

    @staticmethod
    def merge_files(path, target):
        """Write the VRT ``target`` over all ``*.tif`` files in the directory ``path``."""
        directory, vrt = utils.safe_abspath(path), utils.safe_abspath(target)
        if os.path.exists(vrt):
            raise IOError("Target '{}' already exists".format(vrt))
        tiles = glob.glob(os.path.join(directory, "*.tif"))
        if not tiles:
            raise IOError("No source .tif files found in '{}'".format(directory))
        geotiff.write_vrt(vrt, tiles)


def _complete_request(source, request):
    """``to_file`` request with the reference's defaults: the source's projection, the envelope
    of its geometry and its own cell size where the caller gave none."""
    request = dict(request, mode="vals")
    missing = "Cannot determine the {} from the source raster. Please provide {}."
    if "projection" not in request:
        if source.projection is None:
            raise ValueError(missing.format("projection", "a 'projection' argument"))
        request["projection"] = source.projection
    if "bbox" not in request:
        footprint = source.geometry
        if footprint is None:
            raise ValueError(missing.format("extent", "a 'bbox' argument"))
        if hasattr(footprint, "GetEnvelope"):          # an OGR geometry: (x1, x2, y1, y2)
            west, east, south, north = footprint.GetEnvelope()
        else:
            west, south, east, north = footprint.bounds
        request["bbox"] = (west, south, east, north)
    if not ("width" in request and "height" in request):
        grid = source.geo_transform
        if grid is None:
            raise ValueError(missing.format("pixel size", "'width' and 'height' arguments"))
        west, south, east, north = request["bbox"]
        request["width"] = int(round((east - west) / abs(float(grid[1]))))
        request["height"] = int(round((north - south) / abs(float(grid[5]))))
    return request


def to_file(source, url, tile_size, **request):
    """Export ``source`` to disk: GeoTIFF tiles of at most ``tile_size`` cells in a ``tiles``
    directory next to the VRT ``url`` that mosaics them (``url`` may be a directory: the VRT is
    then ``output.vrt`` inside it).  ``bbox``, ``projection``, ``width`` and ``height`` default to
    the source's own extent, projection and cell size; ``start`` / ``stop`` select the frame."""
    request = _complete_request(source, request)
    vrt = utils.safe_abspath(url)
    if os.path.isdir(vrt):
        vrt = os.path.join(vrt, "output.vrt")
    tiles = os.path.join(os.path.dirname(vrt), "tiles")
    RasterTiler(RasterFileSink(source, tiles), tile_size).get_data(**request)
    RasterFileSink.merge_files(tiles, vrt)
