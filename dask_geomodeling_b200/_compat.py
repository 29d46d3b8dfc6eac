"""The three pieces of dask the Block runtime relies on.

The reference imports ``dask.base.tokenize``/``get_scheduler``/
``normalize_token``, ``dask.local.get_sync`` and ``dask.config``
(core/graphs.py:9-10, config.py:12).  When dask is installed those are used
unchanged, so tokens, graph keys and schedulers behave exactly as in the
reference.  When it is not (it is absent from the build image), the minimal
stand-ins below provide deterministic md5 tokens, a synchronous graph
executor and a nested configuration dict with the same call signatures.
"""
import contextlib
import datetime
import hashlib
import pickle
import threading
import uuid

import numpy as np

try:  # pragma: no cover - exercised only where dask is installed
    from dask.base import get_scheduler, normalize_token, tokenize
    from dask.local import get_sync
    from dask import config

    HAVE_DASK = True
except ImportError:
    HAVE_DASK = False

    class _Dispatch:
        """Type -> normaliser registry (subset of dask.utils.Dispatch)."""

        def __init__(self):
            self._lookup = {}

        def register(self, types, func=None):
            if not isinstance(types, tuple):
                types = (types,)

            def wrapper(f):
                for t in types:
                    self._lookup[t] = f
                return f

            return wrapper(func) if func is not None else wrapper

        def __call__(self, obj):
            for klass in type(obj).__mro__:
                if klass in self._lookup:
                    return self._lookup[klass](obj)
            method = getattr(obj, "__dask_tokenize__", None)
            if method is not None:
                return method()
            try:
                return ("__pickle__", pickle.dumps(obj, protocol=4))
            except Exception:
                return ("__random__", uuid.uuid4().hex)

    normalize_token = _Dispatch()

    @normalize_token.register((int, float, str, bytes, type(None), bool, complex, type(Ellipsis)))
    def _normalize_scalar(obj):
        return (type(obj).__name__, obj)

    @normalize_token.register((list, tuple))
    def _normalize_seq(obj):
        return (type(obj).__name__, [normalize_token(x) for x in obj])

    @normalize_token.register(dict)
    def _normalize_dict(obj):
        items = [(normalize_token(k), normalize_token(v)) for k, v in obj.items()]
        return ("dict", sorted(items, key=repr))

    @normalize_token.register((set, frozenset))
    def _normalize_set(obj):
        return ("set", sorted((normalize_token(x) for x in obj), key=repr))

    @normalize_token.register(np.ndarray)
    def _normalize_ndarray(obj):
        data = np.ascontiguousarray(obj)
        if data.dtype.hasobject:
            digest = hashlib.md5(pickle.dumps(data.tolist(), protocol=4)).hexdigest()
        else:
            digest = hashlib.md5(data.view(np.uint8).reshape(-1).data).hexdigest()
        return ("ndarray", digest, str(obj.dtype), obj.shape)

    @normalize_token.register(np.generic)
    def _normalize_npscalar(obj):
        return ("npscalar", str(obj.dtype), obj.item() if obj.dtype.kind != "V" else obj.tobytes())

    @normalize_token.register(np.dtype)
    def _normalize_dtype(obj):
        return ("dtype", str(obj))

    @normalize_token.register((datetime.date, datetime.time, datetime.timedelta, datetime.datetime))
    def _normalize_datetime(obj):
        return (type(obj).__name__, repr(obj))

    @normalize_token.register(type)
    def _normalize_type(obj):
        return ("type", obj.__module__, obj.__qualname__)

    def tokenize(*args, **kwargs):
        """Deterministic md5 token of the arguments."""
        payload = normalize_token(args)
        if kwargs:
            payload = (payload, normalize_token(kwargs))
        return hashlib.md5(repr(payload).encode("utf-8")).hexdigest()

    def get_scheduler(*args, **kwargs):
        getter = config.get("scheduler", None)
        return getter if callable(getter) else None

    def _is_task(x):
        return type(x) is tuple and len(x) > 0 and callable(x[0])

    def get_sync(dsk, keys, pack_exception=None, **kwargs):
        """Evaluate graph keys depth-first in the calling thread."""
        cache = {}

        def resolve(arg):
            if _is_task(arg):
                return arg[0](*[resolve(a) for a in arg[1:]])
            if type(arg) is list:
                return [resolve(a) for a in arg]
            try:
                known = arg in dsk
            except TypeError:  # unhashable literal
                return arg
            return evaluate(arg) if known else arg

        def evaluate(key):
            if key in cache:
                return cache[key]
            task = dsk[key]
            try:
                result = resolve(task) if (_is_task(task) or type(task) is list) else (
                    evaluate(task) if _hashable_key(task, dsk) and task != key else task
                )
            except Exception as e:
                if pack_exception is not None and not getattr(e, "_gm_packed", False):
                    try:
                        pack_exception(e, None)
                    except Exception as packed:
                        packed._gm_packed = True
                        raise
                raise
            cache[key] = result
            return result

        # `resolve` and `evaluate` reference each other, so `cache` sits on a reference cycle:
        # empty it explicitly, otherwise every intermediate raster (HBM) and every pinned
        # result would stay alive until the cyclic garbage collector happens to run
        try:
            if isinstance(keys, list):
                return tuple(evaluate(k) for k in keys)
            return evaluate(keys)
        finally:
            cache.clear()

    def _hashable_key(x, dsk):
        try:
            return x in dsk
        except TypeError:
            return False

    class _Config:
        """Nested dict with dotted-key access (subset of dask.config)."""

        def __init__(self):
            self._data = {}
            self._lock = threading.RLock()

        @staticmethod
        def _merge(base, new, keep_existing):
            for k, v in new.items():
                if isinstance(v, dict) and isinstance(base.get(k), dict):
                    _Config._merge(base[k], v, keep_existing)
                elif not (keep_existing and k in base):
                    base[k] = v

        def update_defaults(self, new):
            with self._lock:
                self._merge(self._data, new, keep_existing=True)

        def get(self, key, default=KeyError):
            node = self._data
            for part in key.split("."):
                if isinstance(node, dict) and part in node:
                    node = node[part]
                else:
                    if default is KeyError:
                        raise KeyError(key)
                    return default
            return node

        def _assign(self, key, value):
            parts = key.split(".")
            node = self._data
            for part in parts[:-1]:
                node = node.setdefault(part, {})
            old = node.get(parts[-1], KeyError)
            node[parts[-1]] = value
            return old

        def set(self, arg=None, **kwargs):
            updates = dict(arg or {})
            updates.update(kwargs)
            with self._lock:
                previous = [(k, self._assign(k, v)) for k, v in updates.items()]
            return _ConfigContext(self, previous)

    class _ConfigContext(contextlib.AbstractContextManager):
        def __init__(self, cfg, previous):
            self._cfg = cfg
            self._previous = previous

        def __exit__(self, *exc):
            for key, old in reversed(self._previous):
                if old is KeyError:
                    parts = key.split(".")
                    node = self._cfg._data
                    for part in parts[:-1]:
                        node = node.get(part, {})
                    node.pop(parts[-1], None)
                else:
                    self._cfg._assign(key, old)
            return False

    config = _Config()
