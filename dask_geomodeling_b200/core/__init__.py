"""Graph runtime: Block base class, compute-graph construction and the fusion pass."""
from . import graphs
from .graphs import Block, DummyBlock, compute, construct, construct_multiple

__all__ = list(graphs.__all__)
