"""Host-side logic of the GPU path that needs no device: polygon soups, stripe ownership of
polygons, request tiling, program-cache keys."""
import numpy as np
import pytest

from dask_geomodeling_b200 import parallel, raster, utils, workloads
from dask_geomodeling_b200.core import fusion


def polygons():
    shell = [(1, 1), (9, 1), (9, 9), (1, 9)]
    hole = [(2, 2), (3, 2), (3, 3)]
    return [
        utils.Polygon([(0, 0), (4, 0), (4, 4)]),
        utils.Polygon(shell, [hole]),
        utils.MultiPolygon([utils.Polygon([(5, 5), (6, 5), (6, 7)]), utils.Polygon([(20, 1), (22, 1), (22, 3)])]),
        utils.Polygon([(30, 30), (31, 30), (31, 35)]),
    ]


def test_expand_ranges():
    np.testing.assert_array_equal(utils._expand_ranges([5, 0, 10], [3, 0, 2]), [5, 6, 7, 10, 11])
    assert len(utils._expand_ranges([], [])) == 0
    assert len(utils._expand_ranges([3, 9], [0, 0])) == 0


def test_soup_subset_and_bounds():
    geoms = polygons()
    soup = utils.PolygonSoup(geoms)
    np.testing.assert_array_equal(soup.bounds(), [[0, 0, 4, 4], [1, 1, 9, 9], [5, 1, 22, 7], [30, 30, 31, 35]])
    for ids in ([2, 1], [3], [0, 1, 2, 3], []):
        sub = soup.subset(ids)
        ref = utils.PolygonSoup([geoms[i] for i in ids])
        np.testing.assert_array_equal(sub.xy, ref.xy.reshape(-1, 2))
        np.testing.assert_array_equal(sub.ring_offsets, ref.ring_offsets)
        np.testing.assert_array_equal(sub.poly_offsets, ref.poly_offsets)


def test_stripe_ownership_of_polygons():
    # 100 rows in 4 stripes of 25; y runs north (row 0) to south
    height = 100
    def square(row0, row1):   # rows [row0, row1) -> y in (height - row1, height - row0)
        return utils.Polygon([(10, height - row1 + 0.3), (20, height - row1 + 0.3),
                              (20, height - row0 - 0.3), (10, height - row0 - 0.3)])
    geoms = [square(2, 10),      # inside stripe 0
             square(20, 30),     # crosses 0 -> 1 (near)
             square(26, 49),     # inside stripe 1 (with the one-row slack still inside)
             square(40, 80),     # reaches stripes 1, 2, 3: far
             square(90, 99),     # inside the last stripe
             square(-30, -10)]   # outside the raster
    soup = utils.PolygonSoup(geoms)
    owner, near, far, crosses, halo, keep = parallel._stripe_ownership(soup, (0, 0, 50, height), height, 4)
    np.testing.assert_array_equal(owner[:5], [0, 0, 1, 1, 3])
    np.testing.assert_array_equal(near, [False, True, False, False, False, False])
    np.testing.assert_array_equal(far, [False, False, False, True, False, False])
    assert not crosses[5]
    # stripe 0 needs the rows of polygon 1 below row 25 (+ slack), and keeps its own rows from 19 up
    assert halo[0] >= 5 and halo[0] <= 7 and keep[0] >= 5 and keep[0] <= 7
    assert halo[1] == 0 and halo[2] == 0 and halo[3] == 0
    # cached per (bbox, height, world)
    assert parallel._stripe_ownership(soup, (0, 0, 50, height), height, 4)[0] is owner
    assert parallel._stripe_ownership(soup, (0, 0, 50, height), height, 2)[0] is not owner


def test_raster_tiler_cuts_the_request_on_its_cell_grid():
    a, b = workloads.cfg1_arrays(64)
    view = workloads.cfg1_view(a, b)
    tiler = raster.RasterTiler(view, [25, 40])
    request = workloads.request(64, 64)
    request.update(bbox=(4, 8, 64, 60), width=60, height=52)
    plan = tiler.get_sources_and_requests(**request)
    kwargs, tiles = plan[0][0], [r for _, r in plan[1:]]
    assert kwargs["shape_yx"] == (52, 60) and kwargs["placements"] == [(0, 0), (0, 25), (0, 50), (40, 0), (40, 25), (40, 50)]
    covered = np.zeros((52, 60), dtype=int)
    for (row0, col0), tile in zip(kwargs["placements"], tiles):
        assert tile["width"] <= 25 and tile["height"] <= 40
        x1, y1, x2, y2 = tile["bbox"]
        assert (x2 - x1) == tile["width"] and (y2 - y1) == tile["height"]          # cell size 1 kept
        assert x1 == 4 + col0 and y2 == 60 - row0
        covered[row0:row0 + tile["height"], col0:col0 + tile["width"]] += 1
    assert (covered == 1).all()
    # non-vals requests and point requests pass through
    assert tiler.get_sources_and_requests(**dict(request, mode="meta"))[0] == (None, None)
    point = dict(request, bbox=(5, 5, 5, 5), width=1, height=1)
    assert tiler.get_sources_and_requests(**point)[0] == (None, None)
    for bad in ([1, 2, 3], 0, [4, -1]):
        with pytest.raises(ValueError):
            raster.RasterTiler(view, bad)
    with pytest.raises(TypeError):
        raster.RasterTiler("not a block", 16)
    # stitching (no device involved): tiles are placed by their offsets, missing tiles keep the fill
    stitched = raster.RasterTiler.process(
        {"dtype": "u1", "fillvalue": 255, "shape_yx": (3, 4), "placements": [(0, 0), (0, 2), (2, 0)]},
        {"values": np.full((1, 2, 2), 1, "u1"), "no_data_value": 255}, None,
        {"values": np.full((1, 1, 4), 3, "u1"), "no_data_value": 255})
    np.testing.assert_array_equal(stitched["values"][0], [[1, 1, 255, 255], [1, 1, 255, 255], [3, 3, 3, 3]])


def test_plan_keys_identify_fused_groups():
    a, b = workloads.cfg1_arrays(16)
    graph, name = workloads.cfg1_view(a, b).get_compute_graph(**workloads.request(16, 16))
    fused = fusion.optimize(graph, name)
    plan = fused[name][1]
    leaves = [(a, workloads.F32_MAX), (b, workloads.F32_MAX)]
    key = fusion._plan_key(plan, leaves)
    assert key is not None and hash(key) == hash(fusion._plan_key(plan, leaves))
    other = fusion._plan_key(plan, [(a.astype("f8"), workloads.F32_MAX), (b, workloads.F32_MAX)])
    assert other != key                                                  # leaf dtypes are part of the key
    graph2, name2 = workloads.cfg1_view(a, a).get_compute_graph(**workloads.request(16, 16))
    assert fusion._plan_key(fusion.optimize(graph2, name2)[name2][1], leaves) != key
    assert fusion._freeze(np.arange(3)) == fusion._freeze(np.arange(3)) != fusion._freeze(np.arange(4))
    assert fusion._freeze(float("nan")) == fusion._freeze(float("nan"))


def test_fused_group_with_an_ndarray_operand():
    """BaseMath accepts ndarray operands; inside a fused group they become extra
    leaves (the lowering used to raise KeyError)."""
    from dask_geomodeling_b200.raster import _program

    a, _ = workloads.cfg1_arrays(8)
    src = workloads.source(a, workloads.F32_MAX)
    ones = np.ones((1, 8, 8), "f4")
    view = raster.Add(raster.Multiply(src, 2.0), ones)
    graph, name = view.get_compute_graph(**workloads.request(8, 8))
    fused = fusion.optimize(graph, name)
    assert fused[name][0] in (fusion.fused_process, fusion.streamed_fused_process)
    plan = fused[name][1]
    with pytest.raises(_program.FusionLimit):
        fusion.build_expression(plan)                    # no place to put the array
    extra = []
    expression = fusion.build_expression(plan, 1, extra)
    assert len(extra) == 1 and extra[0] is ones
    prog, _, results = _program.compile_expression(
        [expression], [(np.dtype("f4"), workloads.F32_MAX), (ones.dtype, None)])
    assert prog.n_inputs == 2 and results[0].dtype == np.float32
    # bins of Classify stay literals (they are not raster operands)
    view = raster.Classify(raster.Multiply(src, 2.0), np.array([10.0, 20.0]))
    graph, name = view.get_compute_graph(**workloads.request(8, 8))
    extra = []
    fusion.build_expression(fusion.optimize(graph, name)[name][1], 1, extra)
    assert extra == []


def test_source_window_covers_what_a_request_reads():
    """`raster.sources.source_window` (the part of a file that RasterFileSource decodes): every
    source cell the nearest-neighbour formula addresses inside the source lies in the window, which
    never reaches beyond the first / last addressed cell by more than the clipping to the source."""
    from dask_geomodeling_b200.raster.sources import source_window, window_geometry

    rng = np.random.default_rng(12)
    src_h, src_w = 600, 530
    gt = utils.GeoTransform((1000.0, 2.5, 0, 2000.0, 0, -2.5))
    for _ in range(200):
        x1, y1 = rng.uniform(700, 2400), rng.uniform(300, 2100)
        w, h = rng.uniform(3, 900), rng.uniform(3, 900)
        width, height = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        bbox = (x1, y1, x1 + w, y1 + h)
        r_lo, r_hi, c_lo, c_hi = source_window(gt, bbox, height, width, src_h, src_w)
        col0, col_step, row0, row_step = window_geometry(gt, bbox, height, width)
        cols = np.floor(col0 + np.arange(width) * col_step).astype(int)
        rows = np.floor(row0 + np.arange(height) * row_step).astype(int)
        cols, rows = cols[(cols >= 0) & (cols < src_w)], rows[(rows >= 0) & (rows < src_h)]
        assert 0 <= r_lo <= r_hi <= src_h and 0 <= c_lo <= c_hi <= src_w
        if len(cols):
            assert c_lo <= cols.min() and cols.max() < c_hi
        if len(rows):
            assert r_lo <= rows.min() and rows.max() < r_hi
        span_c = np.floor([col0, col0 + (width - 1) * col_step]).astype(int)
        span_r = np.floor([row0, row0 + (height - 1) * row_step]).astype(int)
        assert c_lo >= min(max(span_c.min(), 0), src_w) and c_hi <= max(min(span_c.max() + 1, src_w), c_lo)
        assert r_lo >= min(max(span_r.min(), 0), src_h) and r_hi <= max(min(span_r.max() + 1, src_h), r_lo)


def test_prepared_geometry_cache_notices_replaced_elements():
    from dask_geomodeling_b200.geometry import sources

    polygons = [[(i, 0.0), (i + 1.0, 0.0), (i + 1.0, 1.0)] for i in range(100)]
    first = sources._prepared(polygons)
    assert sources._prepared(polygons)[1] is first[1]            # same list: same soup
    polygons[0] = [(50.0, 50.0), (51.0, 50.0), (51.0, 51.0)]       # an end element is always sampled
    again = sources._prepared(polygons)
    assert again[1] is not first[1] and again[2][0].tolist() == [50.0, 50.0, 51.0, 51.0]
