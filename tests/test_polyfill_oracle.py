"""The GDAL scanline-fill restatement (oracle/polyfill.c) against every label pattern
the reference's own tests pin at the GDAL boundary -- CPU only."""
import numpy as np
import pytest

from oracle import polyfill


def box(x1, y1, x2, y2):
    return [[(x1, y1), (x2, y1), (x2, y2), (x1, y2)]]


def test_two_boxes_10x10():
    # reference tests/test_utils.py:336-355
    values, nodata = polyfill.rasterize([box(2, 2, 4, 4), box(6, 6, 8, 8)], (0, 0, 10, 10), 10, 10)
    assert values.dtype == bool and nodata is None
    assert values.sum() == 8
    assert values[0, 6:8, 2:4].all() and values[0, 2:4, 6:8].all()  # row 0 = north


def test_vals_request_2x3():
    # reference tests/test_raster.py:1649-1671
    squares = [
        [[(0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 1.0)]],
        [[(10.0, 2.0), (10.0, 3.0), (20.0, 3.0), (20.0, 2.0)]],
        [[(1.0, 2.0), (1.0, 13.0), (12.0, 13.0), (12.0, 2.0)]],
    ]
    values, nodata = polyfill.rasterize(squares, (0, 0, 2, 3), 3, 2, values=[51, 212, 512])
    flipped = values[0, ::-1]
    assert flipped[1, 0] == 51 and flipped[2, 1] == 512
    assert (flipped == nodata).sum() == 4


def test_overlapping_last_on_top():
    # reference tests/test_raster.py:1673-1683
    squares = [
        [[(0.0, 0.0), (2.0, 0.0), (2.0, 3.0), (0.0, 3.0)]],
        [[(0.0, 1.0), (0.0, 2.0), (1.0, 2.0), (1.0, 1.0)]],
    ]
    values, _ = polyfill.rasterize(squares, (0, 0, 2, 3), 3, 2, values=[0, 1])
    assert values[0, 1, 0] == 1 and (values[0] == 0).sum() == 5


@pytest.mark.parametrize("offset", [0.0, 0.49, 0.51, 1.0])
def test_shifting_pixel(offset):
    # reference tests/test_raster.py:1685-1711: the cell whose centre is covered burns
    pixel = np.array(((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)))
    values, _ = polyfill.rasterize([[pixel + [offset, 0.0]]], (0, 0, 2, 3), 3, 2, values=[0])
    assert values[0, 2, 0 if offset < 0.5 else 1] == 0 and (values == 0).sum() == 1
    values, _ = polyfill.rasterize([[pixel + [0.0, offset]]], (0, 0, 2, 3), 3, 2, values=[0])
    assert values[0, 2 if offset < 0.5 else 1, 0] == 0 and (values == 0).sum() == 1


def test_wkt_rectangle_4x6():
    # reference tests/test_raster_misc.py:258-276: 4 x 6 cells of 0.5
    bbox = (135000, 455997, 135002, 456000)
    values, _ = polyfill.rasterize([box(135000.5, 455998, 135001.5, 455999.5)], bbox, 6, 4)
    assert values[0].astype(int).tolist() == [
        [0, 0, 0, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 1, 1, 0], [0, 0, 0, 0], [0, 0, 0, 0]]


@pytest.mark.parametrize("dx", [0.0, 0.1, 0.4999, 0.50001, 0.9, 0.99999])
def test_no_interaction(dx):
    # reference tests/test_aggregate_raster.py:537-554: second box covers columns 3, 4
    polygons = [box(2.0 + dx, 2.0, 4.0 + dx, 4.0), box(3.0, 6.0, 5, 8.0)]
    labels = polyfill.burn_index(polygons, (0, 0, 10, 10), 10, 10, unlabelled=-1)
    cols = np.nonzero((labels == 1).any(axis=0))[0]
    assert cols.tolist() == [3, 4]


@pytest.mark.parametrize("triangle,covered", [
    ([(2, 2), (1.9, 2), (2, 1.9)], 0), ([(2, 2), (2.1, 2), (2, 1.9)], 0),
    ([(2, 2), (2.1, 2), (2, 2.1)], 0), ([(2, 2), (1.9, 2), (2, 2.1)], 0)])
def test_tiny_triangles_cover_no_centre(triangle, covered):
    # reference tests/test_aggregate_raster.py:568-587: centroid fallback is needed
    labels = polyfill.burn_index([[triangle]], (0, 0, 6, 4), 2, 3, unlabelled=-1)
    assert (labels >= 0).sum() == covered


def test_hole_and_multipolygon_even_odd():
    shell = [(1, 1), (9, 1), (9, 9), (1, 9)]
    hole = [(3, 3), (7, 3), (7, 7), (3, 7)]
    labels = polyfill.burn_index([[shell, hole]], (0, 0, 10, 10), 10, 10, unlabelled=-1)
    assert (labels >= 0).sum() == 64 - 16
    spans = polyfill.spans([[shell, hole]], (0, 0, 10, 10), 10, 10)
    assert (spans[:, 3] - spans[:, 2] + 1).sum() == 48
