"""GeometryBlock contract (reference: geometry/base.py:10-96).

Requests: ``mode`` ('intersects' | 'centroid' | 'extent'), ``geometry``,
``projection``, ``limit``, ``min_size``, ``start``/``stop``, ``filters``.
Responses: ``{"features": DataFrame-with-geometry-column, "projection"}`` or
``{"extent": (x1, y1, x2, y2) | None, "projection"}``.
"""
from ..core import Block

__all__ = ["GeometryBlock"]


class GeometryBlock(Block):
    """Base of all geometry views; subclasses define ``columns``."""

    def __getitem__(self, name):
        raise NotImplementedError("series blocks are outside the CUDA raster path")
