"""Block runtime: lazy views that expand into dask-style compute graphs.

Drop-in for the reference's ``dask_geomodeling.core.graphs`` (core/graphs.py):
same public names (``Block``, ``DummyBlock``, ``construct``,
``construct_multiple``, ``compute``), same graph/key conventions
(``<classname>_<md5 token>``, core/graphs.py:161-190, :220-222), same
serialisation format (:192-286).  What differs is what happens in
``compute``: before the graph is handed to the scheduler, contiguous
element-wise sub-graphs are fused into single CUDA launches
(see ``fusion.py``) and intermediate rasters stay in HBM.
"""
import importlib
import inspect
import json
import logging
from datetime import datetime, timedelta

from .._compat import get_scheduler, get_sync, normalize_token, tokenize

logger = logging.getLogger(__name__)

__all__ = ["construct", "construct_multiple", "compute", "Block", "DummyBlock"]


def _prefix_key_on_error(e, dumps):
    """``pack_exception`` hook: put the failing graph key in front of the message
    (reference behaviour: core/graphs.py:21-27)."""
    key = inspect.currentframe().f_back.f_locals.get("key")
    e.args = ("{0}: {1}".format(key, str(e)),)
    raise e


def _token_from_key(key):
    """'Name_<32 hex chars>' -> the token, else None (core/graphs.py:30-39)."""
    head, sep, tail = key.rpartition("_")
    if not sep or not head or len(tail) != 32:
        return None
    try:
        int(tail, 16)
    except ValueError:
        return None
    return tail.lower()


def compute(graph, name, *args, **kwargs):
    """Evaluate ``graph[name]`` with the configured dask scheduler (synchronous
    when none is configured, as core/graphs.py:42-49).

    Element-wise sub-graphs are fused and rasters are kept on the device
    between tasks; the value returned for ``name`` is always host data.
    """
    from . import fusion

    scheduler = get_scheduler()
    if scheduler is None:
        scheduler = get_sync
    graph = fusion.optimize(graph, name)
    with fusion.device_resident():
        result = scheduler(graph, [name])[0]
        return fusion.to_host(result)


def construct(graph, name, validate=True):
    """Build the Block named ``name`` (and everything it depends on)."""
    return construct_multiple(graph, [name], validate)[0]


def construct_multiple(graph, names, validate=True):
    """Build several Blocks that may share dependencies from a class graph
    ``{key: (BlockClass | 'import.path', *args)}``."""
    factory_graph = {}
    for key, spec in graph.items():
        klass = spec[0]
        if isinstance(klass, str):
            klass = Block.from_import_path(klass)
        if not (inspect.isclass(klass) and issubclass(klass, Block)):
            raise TypeError("Cannot construct from object of type '{}'".format(klass))
        rest = tuple(spec[1:])
        if validate:
            factory_graph[key] = (klass,) + rest
            continue
        token = _token_from_key(key)
        if token is None:
            logger.warning(
                "Construct received a key with an invalid name ('%s'),"
                "while validation was turned off",
                key,
            )
        factory_graph[key] = (klass._init_no_validation, token) + rest
    return get_sync(factory_graph, names, pack_exception=_prefix_key_on_error)


class Block(object):
    """Base of every view node.

    A Block stores its constructor arguments in ``self.args``.  For a request
    it decides which of those arguments must be evaluated, and with what
    request (``get_sources_and_requests``); the evaluated data is then passed
    to the static ``process`` function.
    """

    JSON_VERSION = 2

    def __init__(self, *args):
        self.args = args

    @classmethod
    def _init_no_validation(cls, token, *args):
        block = cls.__new__(cls)
        block.args = args
        if token:
            block._cached_token = token
        return block

    # -- identity ---------------------------------------------------------
    @property
    def token(self):
        """Deterministic hash of class path and arguments, cached per object."""
        cached = self.__dict__.get("_cached_token")
        if cached is None:
            parts = [a.token if isinstance(a, Block) else a for a in self.args]
            cached = self._cached_token = tokenize(self.get_import_path(), *parts)
        return cached

    @property
    def name(self):
        return "{}_{}".format(type(self).__name__, self.token)

    # -- to be overridden -----------------------------------------------------
    @staticmethod  # graph tuples store the function itself: keep it static
    def process(data):
        return data

    def get_sources_and_requests(self, **request):
        """Iterable of ``(source, request)``; non-Block sources are passed to
        ``process`` literally and ``request=None`` suppresses recursion."""
        return ((source, request) for source in self.args)

    # -- evaluation -------------------------------------------------------------
    def get_data(self, **request):
        return compute(*self.get_compute_graph(**request))

    def get_compute_graph(self, cached_compute_graph=None, **request):
        """``(graph, name)`` with ``graph[name] = (process, *args)``; string args
        that are keys of ``graph`` refer to the outputs of other tasks."""
        name = "{}_{}".format(type(self).__name__.lower(), tokenize([self.token, request]))
        graph = cached_compute_graph or dict()
        if name in graph:
            return graph, name
        task = [self.process]
        for source, source_request in self.get_sources_and_requests(**request):
            if isinstance(source, Block) and source_request is not None:
                graph, key = source.get_compute_graph(cached_compute_graph=graph, **source_request)
                task.append(key)
            else:
                task.append(source)
        graph[name] = tuple(task)
        return graph, name

    # -- (de)serialisation ----------------------------------------------------------
    def get_graph(self, serialize=False):
        """Class graph ``{name: [cls-or-path, *args]}`` describing this view."""
        graph = {}
        spec = [self.get_import_path() if serialize else type(self)]
        for arg in self.args:
            if isinstance(arg, Block):
                sub, key = arg.get_graph(serialize=serialize)
                graph.update(sub)
                spec.append(key)
            else:
                spec.append(arg)
        graph[self.name] = spec
        return graph, self.name

    def __reduce__(self):
        return construct, self.get_graph() + (False,)

    @classmethod
    def get_import_path(cls):
        module, name = cls.__module__, cls.__name__
        try:
            found = getattr(importlib.import_module(module), name)
        except (ImportError, KeyError, AttributeError):
            raise Exception("Can't serialize %r: it's not found as %s.%s" % (cls, module, name))
        if found is not cls:
            raise Exception(
                "Can't serialize %r: it's not the same object as %s.%s" % (cls, module, name)
            )
        return "{}.{}".format(module, name)

    @staticmethod
    def from_import_path(path):
        module, name = path.rsplit(".", 1)
        klass = getattr(importlib.import_module(module), name)
        if inspect.isclass(klass) and issubclass(klass, Block):
            return klass
        raise TypeError('"{}" is not valid Block.'.format(path))

    def serialize(self):
        graph, name = self.get_graph(serialize=True)
        return {"version": self.JSON_VERSION, "graph": graph, "name": name}

    @classmethod
    def deserialize(cls, val, validate=False):
        return construct(val["graph"], val["name"], validate=validate)

    def to_json(self, **kwargs):
        return json.dumps(self.serialize(), **kwargs)

    @classmethod
    def from_json(cls, val, **kwargs):
        return cls.deserialize(json.loads(val, **kwargs))

    def __repr__(self):
        return "{}({})".format(type(self).__name__, ", ".join(repr(x) for x in self.args))


class DummyBlock(Block):
    """Stands in for a block of which only the name (and thus token) is known."""

    def __init__(self, name):
        super().__init__(name)

    @property
    def name(self):
        return self.args[0]

    @property
    def token(self):
        return self.name.split("_")[1]


@normalize_token.register((datetime, timedelta))
def _normalize_datetime(value):
    return hash(value)


try:  # shapely is optional in this build
    from shapely.geometry.base import BaseGeometry

    @normalize_token.register(BaseGeometry)
    def _normalize_shapely(geometry):
        return geometry.wkb
except ImportError:  # pragma: no cover
    pass
