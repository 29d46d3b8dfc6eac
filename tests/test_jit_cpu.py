"""The specialiser without a GPU: every generated kernel is valid sm_100a code (NVRTC compiles it
here), and the source of a program does not depend on its scalar constants (they travel in the
kernel's parameter block), so that a changed constant reuses the compiled kernel."""
import ctypes

import numpy as np
import pytest

from dask_geomodeling_b200 import _native, workloads
from dask_geomodeling_b200.raster import _program
from dask_geomodeling_b200.raster._program import Leaf, Node


def _dtypes(leaf_types, results):
    ins = (ctypes.c_int32 * max(len(leaf_types), 1))(*[_native.dtype_code(d) for d, _ in leaf_types])
    outs = (ctypes.c_int32 * len(results))(*[_native.dtype_code(t.dtype) for t in results])
    return ins, outs


def source_of(node, leaf_types):
    prog, comp, results = _program.compile_expression([node], leaf_types)
    ins, outs = _dtypes(leaf_types, results)
    lib = _native.load_library()
    length = ctypes.c_int64()
    assert lib.gm_jit_source(ctypes.byref(prog), ins, outs, None, 0, ctypes.byref(length)) == 0
    buf = ctypes.create_string_buffer(length.value + 1)
    assert lib.gm_jit_source(ctypes.byref(prog), ins, outs, buf, length.value + 1, ctypes.byref(length)) == 0
    return buf.value.decode(), (prog, comp, ins, outs)


def nvrtc_available():
    import glob

    return bool(glob.glob("/usr/local/cuda/lib64/libnvrtc.so*"))


F4 = (np.dtype("f4"), workloads.F32_MAX)


def chain(scale, threshold):
    p = Node("multiply", [Node("add", [Leaf(0), Leaf(1)], dtype="float32", fillvalue=workloads.F32_MAX), scale],
             dtype="float32", fillvalue=workloads.F32_MAX)
    return Node("maskbelow", [p], value=threshold)


def test_source_is_independent_of_the_scalars():
    a, _ = source_of(chain(0.5, 40.0), [F4, F4])
    b, _ = source_of(chain(0.75, 12.5), [F4, F4])
    assert a == b
    assert "p.k[" in a                                  # constants are kernel parameters
    c, _ = source_of(Node("maskbelow", [Node("add", [Leaf(0), Leaf(1)], dtype="float32",
                                             fillvalue=workloads.F32_MAX)], value=1.0), [F4, F4])
    assert c != a          # another structure is another kernel


@pytest.mark.skipif(not nvrtc_available(), reason="no NVRTC in this environment")
@pytest.mark.parametrize("name", ["chain", "reduce_max", "reduce_count_u1", "cfg2"])
def test_generated_kernels_compile_for_sm_100a(name):
    lib = _native.load_library()
    if name == "chain":
        node, leaves = chain(0.5, 40.0), [F4, F4]
    elif name == "reduce_max":
        node = Node("reduce", [Leaf(0), Leaf(1), Leaf(2)], statistic="max", dtype="float32",
                    fillvalue=workloads.F32_MAX)
        leaves = [F4, F4, (np.dtype("u1"), 255)]
    elif name == "reduce_count_u1":
        node = Node("reduce", [Leaf(0), Leaf(1)], statistic="count", dtype="uint8", fillvalue=255)
        leaves = [(np.dtype("u1"), 255), (np.dtype("u1"), 7)]
    else:
        r = Node("reclassify", [Leaf(0)], dtype="int64", fillvalue=np.iinfo("i8").max,
                 data=workloads.CFG2_PAIRS, select=True)
        st = Node("step", [Node("clip", [Leaf(1), r])], left=0, right=1, value=50.0, at=0.5)
        node, leaves = Node("isdata", [st]), [(np.dtype("i2"), 32767), F4]
    _, (prog, comp, ins, outs) = source_of(node, leaves)
    size = ctypes.c_int64()
    rc = lib.gm_jit_check(ctypes.byref(prog), ins, outs, ctypes.byref(size))
    assert rc == 0, lib.gm_last_error().decode()
    assert size.value > 1000
