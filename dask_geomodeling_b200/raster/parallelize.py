"""Blocks that parallelize raster blocks (reference: raster/parallelize.py).

``RasterTiler`` cuts a ``vals`` request in tiles of at most ``tile_size`` cells; every tile
is an independent sub-graph (on the GPU path: its own upload, launches and download, which
the graph runtime may overlap), and ``process`` stitches the tiles.  Row stripes over the
GPUs of one box are the same idea across processes (``dask_geomodeling_b200/parallel.py``).
"""
import numpy as np

from .. import _native
from .base import BaseSingle

__all__ = ["RasterTiler"]


class RasterTiler(BaseSingle):
    """Parallelize operations on a RasterBlock by tiling the request.

    Args:
      source (RasterBlock): the source raster
      tile_size (int or list): maximum size of a tile in cells; ``[width, height]`` to give
        different sizes for the two directions.
    """

    def __init__(self, source, tile_size):
        if hasattr(tile_size, "__iter__"):
            if len(tile_size) != 2:
                raise ValueError("'tile_size' should be a scalar or a list of length 2.")
            tile_size = [int(x) for x in tile_size]
        else:
            tile_size = [int(tile_size), int(tile_size)]
        if tile_size[0] <= 0 or tile_size[1] <= 0:
            raise ValueError("'tile_size' should be greater than 0")
        super().__init__(source, tile_size)

    @property
    def tile_size(self):
        return self.args[1]

    def get_sources_and_requests(self, **request):
        if request["mode"] != "vals":
            return [(None, None), (self.store, request)]
        x1, y1, x2, y2 = request["bbox"]
        width, height = request["width"], request["height"]
        if x1 == x2 and y1 == y2:
            return [(None, None), (self.store, request)]   # point requests pass through
        # tiles are cut on the CELL grid of the request (row 0 = north), so that every tile
        # request addresses exactly the cells it will fill
        tile_w, tile_h = self.tile_size
        cell_x, cell_y = (x2 - x1) / width, (y2 - y1) / height
        placements, requests = [], []
        for row0 in range(0, height, tile_h):
            row1 = min(row0 + tile_h, height)
            for col0 in range(0, width, tile_w):
                col1 = min(col0 + tile_w, width)
                tile = dict(request)
                tile["bbox"] = (x1 + col0 * cell_x, y2 - row1 * cell_y, x1 + col1 * cell_x, y2 - row0 * cell_y)
                tile["width"], tile["height"] = col1 - col0, row1 - row0
                placements.append((row0, col0))
                requests.append((self.store, tile))
        kwargs = {"dtype": self.dtype, "fillvalue": self.fillvalue, "shape_yx": (height, width),
                  "placements": placements}
        return [(kwargs, None)] + requests

    @staticmethod
    def process(process_kwargs, *all_data):
        if len(all_data) == 0:
            return None
        if process_kwargs is None:
            return all_data[0]   # non-tiled / meta / time requests
        filled = [d for d in all_data if d is not None]
        if not filled:
            return None
        bands = filled[0]["values"].shape[0]
        fill = process_kwargs["fillvalue"]
        out = np.full((bands,) + tuple(process_kwargs["shape_yx"]), fill, process_kwargs["dtype"])
        for (row0, col0), data in zip(process_kwargs["placements"], all_data):
            if data is None:
                continue   # a tile without data keeps the fill value
            tile = data["values"]
            if _native.is_device(tile):
                tile = np.asarray(tile)
            out[:, row0:row0 + tile.shape[1], col0:col0 + tile.shape[2]] = tile
        return {"values": out, "no_data_value": fill}
