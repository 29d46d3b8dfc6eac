"""Import the REAL reference modules behind stubbed third-party packages.

TEST INFRASTRUCTURE ONLY, and only usable where ``/root/reference`` exists (the
build container).  The reference needs GDAL, shapely, geopandas, pyproj, pytz
and dask at import time; none of them is installed here.  Everything on the
NumPy/SciPy/pandas side of the hot path still runs once those imports are
satisfied by empty stand-ins.  ``load()`` returns a
namespace with the reference modules; it is used by
``tests/golden/make_golden.py`` to generate golden vectors and by
``tests/test_oracle_vs_reference.py`` to pin ``oracle.raster`` directly.
"""
import hashlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dask_geomodeling"))


class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to attribute sinks."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _module(name, **attrs):
    mod = _StubModule(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    """Attribute sink: any attribute access or call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _install_stubs():
    import numpy  # noqa: F401  (real)
    import pandas  # noqa: F401  (must be imported before pytz is stubbed)
    import scipy.ndimage  # noqa: F401

    if "osgeo" not in sys.modules:
        gdal = _module("osgeo.gdal", VersionInfo=lambda *a: "3080400", UseExceptions=lambda: None,
                       SetConfigOption=lambda *a: None, GRA_NearestNeighbour=0, GA_Update=1)
        ogr = _module("osgeo.ogr", UseExceptions=lambda: None, OFTReal=2, OFTInteger=0)
        osr = _module("osgeo.osr", UseExceptions=lambda: None)
        gdal_array = _module("osgeo.gdal_array")
        _module("osgeo", gdal=gdal, ogr=ogr, osr=osr, gdal_array=gdal_array)
    if "shapely" not in sys.modules:
        class BaseGeometry:  # core/graphs.py:12 registers a tokenizer on it
            pass

        geometry = _module("shapely.geometry", box=_Anything(), Point=_Anything, Polygon=_Anything,
                           shape=_Anything())
        base = _module("shapely.geometry.base", BaseGeometry=BaseGeometry)
        geometry.base = base
        _module("shapely.ops", transform=_Anything())
        _module("shapely.errors", ShapelyError=Exception)
        _module("shapely.wkt", loads=_Anything())
        _module("shapely", geometry=geometry, from_wkt=_Anything(), GEOSException=Exception,
                Point=_Anything)
    if "pyproj" not in sys.modules:
        _module("pyproj.exceptions", ProjError=type("ProjError", (Exception,), {}))
        _module("pyproj", CRS=_Anything, Transformer=_Anything)
    for name in ("geopandas", "pyogrio"):
        if name not in sys.modules:
            _module(name, GeoSeries=_Anything, GeoDataFrame=_Anything)
    if "pytz" not in sys.modules:
        _module("pytz", timezone=_Anything(), UTC=None)
    if "dask" not in sys.modules:
        store = {}

        def update_defaults(new):
            for k, v in new.items():
                store.setdefault(k, {}).update(v) if isinstance(v, dict) else store.setdefault(k, v)

        def get(key, default=None):
            node = store
            for part in key.split("."):
                if not isinstance(node, dict) or part not in node:
                    return default
                node = node[part]
            return node

        def set_(arg=None, **kw):
            for k, v in dict(arg or {}, **kw).items():
                parts = k.split(".")
                node = store
                for p in parts[:-1]:
                    node = node.setdefault(p, {})
                node[parts[-1]] = v

        config = _module("dask.config", update_defaults=update_defaults, get=get, set=set_)

        class _Normalize:
            def register(self, types_, func=None):
                return (lambda f: f) if func is None else func

        def tokenize(*args, **kwargs):
            return hashlib.md5(repr((args, kwargs)).encode()).hexdigest()

        def get_sync(dsk, keys, **kwargs):
            def run(key):
                task = dsk[key]
                if isinstance(task, tuple) and callable(task[0]):
                    return task[0](*[run(a) if isinstance(a, str) and a in dsk else a for a in task[1:]])
                return task
            return [run(k) for k in keys] if isinstance(keys, list) else run(keys)

        base = _module("dask.base", tokenize=tokenize, normalize_token=_Normalize(),
                       get_scheduler=lambda *a, **k: None)
        local = _module("dask.local", get_sync=get_sync)
        _module("dask", config=config, base=base, local=local)


_loaded = None


def load():
    """Namespace with the reference's hot-path modules (elemwise, misc, spatial,
    temporal, reduction, combine, aggregate, measurements, utils)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not found at {}".format(REFERENCE_ROOT))
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.utils = importlib.import_module("dask_geomodeling.utils")
    ns.elemwise = importlib.import_module("dask_geomodeling.raster.elemwise")
    ns.misc = importlib.import_module("dask_geomodeling.raster.misc")
    ns.spatial = importlib.import_module("dask_geomodeling.raster.spatial")
    ns.temporal = importlib.import_module("dask_geomodeling.raster.temporal")
    ns.measurements = importlib.import_module("dask_geomodeling.measurements")
    ns.aggregate = importlib.import_module("dask_geomodeling.geometry.aggregate")
    ns.reduction = importlib.import_module("dask_geomodeling.raster.reduction")
    ns.combine = importlib.import_module("dask_geomodeling.raster.combine")
    _loaded = ns
    return ns
