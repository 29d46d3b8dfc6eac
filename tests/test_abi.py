"""The drop-in boundary without a GPU: libgeokernels.so loads, exports every entry point that
include/geokernels.h declares, and the ctypes prototypes of the host side cover them all.
No compute call is made (that is what the ``-m gpu`` tests are for)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "geokernels.h")


@pytest.fixture(scope="module")
def library():
    from dask_geomodeling_b200.csrc import build

    return ctypes.CDLL(build.build())   # compiles only what is out of date


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # comments mention functions too
    return sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for required in ("gm_init", "gm_last_error", "gm_eval_program", "gm_smooth", "gm_moving_max", "gm_dilate",
                     "gm_hillshade", "gm_temporal_aggregate", "gm_temporal_cumulative",
                     "gm_rasterize_polygons", "gm_zonal_stats", "gm_zonal_partials_device",
                     "gm_zonal_finalize_device", "gm_zonal_values", "gm_segment_order_stat", "gm_resample_nn"):
        assert required in names


def test_library_exports_every_declared_symbol(library):
    missing = [name for name in declared_functions() if not hasattr(library, name)]
    assert not missing, "declared in include/geokernels.h but not exported: {}".format(missing)


def test_host_side_prototypes_cover_the_header(library):
    from dask_geomodeling_b200 import _native

    declared = set(declared_functions())
    bound = set(_native.SYMBOLS)
    assert bound <= declared, "bound but not declared: {}".format(sorted(bound - declared))
    # everything the Python host side calls is declared; the header may offer more
    assert _native.load_library().gm_abi_version() == 1


def test_no_cpu_fallback_and_no_oracle_in_the_product():
    # the product path must not import the oracle (test infrastructure only)
    for folder, _, files in os.walk(os.path.join(ROOT, "dask_geomodeling_b200")):
        for name in files:
            if name.endswith(".py"):
                text = open(os.path.join(folder, name)).read()
                assert "import oracle" not in text and "from oracle" not in text, name
