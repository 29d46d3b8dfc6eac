"""Second, independent pin of ``oracle/polyfill.c`` (parity gap 1).

The reference's tests pin GDAL's fill rule only on axis-aligned boxes and tiny triangles
(``tests/test_polyfill_oracle.py``).  Here every slanted-edge result is checked against the
DEFINITION of the rule -- "a cell is burned when its centre is inside the polygon, even-odd
over all rings of the feature" (reference: tests/test_raster.py:1685-1711,
tests/test_aggregate_raster.py:557-587) -- evaluated in exact rational arithmetic
(``fractions.Fraction`` of the float64 vertex coordinates, no rounding anywhere):

* whole grids through an exact scanline (per row: the exact crossings of the row-centre line,
  each turned into the first column whose centre lies right of it);
* sampled cells through a direct crossing-number test of the centre, concentrated on cells
  next to a label change, i.e. next to polygon boundaries.

Vertices stay off ``k + 0.5`` (the reference declares centres exactly on an edge ill-defined,
tests/test_raster.py:1686).  No GPU needed: this pins the checker, not the product.
"""
from fractions import Fraction
from math import floor

import numpy as np
import pytest

from oracle import polyfill

NONE = np.iinfo(np.int32).max


def random_features(n, size, seed, holes=True):
    """Concave star-ish polygons, some with a hole, some multi-part; 3-decimals + 0.0137."""
    rng = np.random.default_rng(seed)
    features = []
    for _ in range(n):
        parts = []
        for _ in range(int(rng.integers(1, 3))):
            cx, cy = rng.uniform(-0.1 * size, 1.1 * size, 2)
            k = int(rng.integers(3, 14))
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            base = rng.uniform(0.02 * size, size / 4)
            rad = base * rng.uniform(0.25, 1.0, k)
            ring = np.round(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1), 3) + 0.0137
            parts.append(ring)
            if holes and rng.random() < 0.4:
                kh = int(rng.integers(3, 8))
                angh = np.sort(rng.uniform(0, 2 * np.pi, kh))
                radh = base * rng.uniform(0.05, 0.22, kh)
                parts.append(np.round(np.stack([cx + radh * np.cos(angh), cy + radh * np.sin(angh)], 1), 3) + 0.0137)
        features.append(parts)
    return features


def exact_edges(rings, bbox, height, width):
    """Edges of all rings in exact pixel coordinates (x right, y down from the top)."""
    x1, y1, x2, y2 = (Fraction(v) for v in bbox)
    dx, dy = (x2 - x1) / width, (y2 - y1) / height
    edges = []
    for ring in rings:
        pts = [((Fraction(float(x)) - x1) / dx, (y2 - Fraction(float(y))) / dy) for x, y in ring]
        if pts[0] != pts[-1]:
            pts.append(pts[0])
        edges.extend(zip(pts[:-1], pts[1:]))
    return edges


def exact_burn(features, bbox, height, width):
    """Labels by exact scanline: later features on top."""
    labels = np.full((height, width), NONE, dtype=np.int32)
    for index, rings in enumerate(features):
        edges = exact_edges(rings, bbox, height, width)
        ys = [p[1] for e in edges for p in e]
        r_lo, r_hi = max(int(floor(min(ys))) - 1, 0), min(int(floor(max(ys))) + 1, height - 1)
        for row in range(r_lo, r_hi + 1):
            cy = Fraction(2 * row + 1, 2)
            parity = np.zeros(width + 1, dtype=np.int64)
            for (ax, ay), (bx, by) in edges:
                if (ay <= cy) == (by <= cy):
                    continue                      # the edge does not cross the centre line
                xi = ax + (cy - ay) * (bx - ax) / (by - ay)
                # first column whose centre lies strictly right of the crossing
                first = floor(xi - Fraction(1, 2)) + 1
                parity[min(max(first, 0), width)] += 1
            inside = (np.cumsum(parity)[:width] & 1).astype(bool)
            labels[row, inside] = index
    return labels


def centre_inside(edges, col, row):
    """Crossing number of the ray from the centre towards -x, exact."""
    cx, cy = Fraction(2 * int(col) + 1, 2), Fraction(2 * int(row) + 1, 2)
    crossings = 0
    for (ax, ay), (bx, by) in edges:
        if (ay <= cy) == (by <= cy):
            continue
        xi = ax + (cy - ay) * (bx - ax) / (by - ay)
        assert xi != cx, "centre exactly on an edge: ill-defined input"
        crossings += xi < cx
    return bool(crossings & 1)


@pytest.mark.parametrize("seed,shape,bbox", [
    (1, (96, 128), (0, 0, 128, 96)),
    (2, (150, 110), (0, 0, 110, 150)),
    (3, (80, 80), (135000.0, 455960.0, 135040.0, 456000.0)),     # cell size 0.5, large offsets
    (4, (64, 200), (-50.0, -16.0, 50.0, 16.0)),                   # cell size 0.5, around zero
])
def test_burn_index_equals_exact_centre_rule(seed, shape, bbox):
    height, width = shape
    x1, y1, x2, y2 = bbox
    unit = random_features(40, 1000.0, seed)                      # in ~[0, 1000]^2, scaled below
    features = []
    for rings in unit:
        scaled = []
        for r in rings:
            xy = np.empty_like(r)
            # 3 decimals of a cell + 0.0137 cells: never on a cell-centre line
            xy[:, 0] = x1 + (np.round(r[:, 0] / 1000.0 * width, 3) + 0.0137) * (x2 - x1) / width
            xy[:, 1] = y1 + (np.round(r[:, 1] / 1000.0 * height, 3) + 0.0137) * (y2 - y1) / height
            scaled.append(xy)
        features.append(scaled)
    got = polyfill.burn_index(features, bbox, height, width)
    expected = exact_burn(features, bbox, height, width)
    np.testing.assert_array_equal(got, expected)
    assert (got != NONE).sum() > 0.2 * got.size                    # the polygons do cover cells


def test_single_features_cell_by_cell_near_boundaries():
    """One feature at a time (no overwrite), direct crossing-number test of the cells next to a
    label change plus a random sample of the others."""
    height, width, bbox = 90, 120, (0, 0, 120, 90)
    rng = np.random.default_rng(9)
    for index, rings in enumerate(random_features(25, 110, seed=11)):
        got = polyfill.burn_index([rings], bbox, height, width) == 0
        edges = exact_edges(rings, bbox, height, width)
        change = np.zeros_like(got)
        change[:, 1:] |= got[:, 1:] != got[:, :-1]
        change[:, :-1] |= got[:, 1:] != got[:, :-1]
        change[1:, :] |= got[1:, :] != got[:-1, :]
        change[:-1, :] |= got[1:, :] != got[:-1, :]
        cells = list(zip(*np.nonzero(change)))
        cells += [(int(r), int(c)) for r, c in zip(rng.integers(0, height, 150), rng.integers(0, width, 150))]
        for row, col in cells:
            assert got[row, col] == centre_inside(edges, col, row), (index, row, col)


def test_holes_and_multipolygons_are_even_odd_over_all_rings():
    shell = [(1.2, 1.2), (58.7, 1.2), (58.7, 58.7), (1.2, 58.7)]
    hole = [(10.2, 10.2), (30.7, 12.4), (28.1, 40.1), (12.9, 37.3)]
    island = [(15.3, 15.3), (25.1, 16.2), (24.4, 30.3), (16.8, 28.9)]     # inside the hole: filled again
    other = [(70.4, 5.2), (95.3, 9.1), (88.8, 40.6)]
    feature = [shell, hole, island, other]
    got = polyfill.burn_index([feature], (0, 0, 100, 60), 60, 100)
    expected = exact_burn([feature], (0, 0, 100, 60), 60, 100)
    np.testing.assert_array_equal(got, expected)
    # row 0 is north: cell (row, col) has its centre at (col + 0.5, 60 - row - 0.5)
    assert got[37, 20] == 0 and got[47, 12] == NONE and got[54, 5] == 0 and got[41, 84] == 0
