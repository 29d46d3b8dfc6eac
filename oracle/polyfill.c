/*
 * CPU restatement of GDAL's scanline polygon fill -- ORACLE, test infrastructure only.
 *
 * The reference burns polygons with gdal.RasterizeLayer (utils.py:748-754) and GDAL
 * is NOT vendored under /root/reference (third-party dependency `osgeo`, CI-pinned to
 * GDAL 3.4.1 / 3.8.4 / 3.10.3 in .github/workflows/test.yml:21-41).  This file restates
 * GDAL's published algorithm, alg/llrasterize.cpp::GDALdllImageFilledPolygon, as used
 * by alg/gdalrasterize.cpp without ALL_TOUCHED:
 *
 *   - vertices go to pixel/line space with the inverse geotransform
 *     (GDALInvGeoTransform, non-rotated case: inv0 = -gt0/gt1, inv1 = 1/gt1,
 *      inv3 = -gt3/gt5, inv5 = 1/gt5;  px = inv0 + x*inv1 + y*0, py = inv3 + x*0 + y*inv5);
 *   - for every row y the scanline sits at dy = y + 0.5;
 *   - an edge (ring-closing edges included) contributes when dy1 <= dy < dy2 (after
 *     ordering its end points by y) at x = floor((dy-dy1)*(dx2-dx1)/(dy2-dy1) + dx1 + 0.5);
 *   - horizontal edges lying exactly on a scanline: only "bottom" ones (x1 > x2) are
 *     filled, separately, over [floor(x2+.5), floor(x1+.5) - 1];
 *   - crossings of ALL rings of the (multi)polygon are sorted together and filled
 *     pairwise [x_i, x_{i+1} - 1] (even-odd), clipped to the raster;
 *   - features are burned in order, later ones overwrite earlier ones.
 *
 * Pinned against the label patterns of the reference's own tests (see
 * tests/test_oracle_golden.py: tests/test_utils.py:349-355, tests/test_raster.py:1663-1711,
 * tests/test_raster_misc.py:258-276, tests/test_aggregate_raster.py:537-577).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int cmp_int(const void* a, const void* b) {
  int x = *(const int*)a, y = *(const int*)b;
  return (x > y) - (x < y);
}

/* callback: burn columns [x0, x1] of row y for polygon p */
typedef void (*span_fn)(void* ctx, int64_t p, int y, int x0, int x1);

static void fill_polygons(const double* xy, const int64_t* ring_offsets, const int64_t* poly_offsets,
                          int64_t n_polygons, const double* gt, int height, int width,
                          span_fn emit, void* ctx) {
  const double inv0 = -gt[0] / gt[1], inv1 = 1.0 / gt[1];
  const double inv3 = -gt[3] / gt[5], inv5 = 1.0 / gt[5];
  for (int64_t p = 0; p < n_polygons; ++p) {
    const int64_t r0 = poly_offsets[p], r1 = poly_offsets[p + 1];
    if (r1 <= r0) continue;
    const int64_t v0 = ring_offsets[r0], v1 = ring_offsets[r1];
    const int64_t n = v1 - v0;
    if (n <= 0) continue;
    double* px = (double*)malloc(sizeof(double) * n);
    double* py = (double*)malloc(sizeof(double) * n);
    int* ints = (int*)malloc(sizeof(int) * (n + 1));
    for (int64_t i = 0; i < n; ++i) {
      const double x = xy[2 * (v0 + i)], y = xy[2 * (v0 + i) + 1];
      px[i] = inv0 + x * inv1 + y * 0.0;
      py[i] = inv3 + x * 0.0 + y * inv5;
    }
    double dminy = py[0], dmaxy = py[0];
    for (int64_t i = 1; i < n; ++i) {
      if (py[i] < dminy) dminy = py[i];
      if (py[i] > dmaxy) dmaxy = py[i];
    }
    int miny = (int)dminy, maxy = (int)dmaxy;
    if (miny < 0) miny = 0;
    if (maxy >= height) maxy = height - 1;
    const int minx = 0, maxx = width - 1;
    for (int y = miny; y <= maxy; ++y) {
      const double dy = y + 0.5;
      int count = 0;
      for (int64_t r = r0; r < r1; ++r) {
        const int64_t a = ring_offsets[r] - v0, b = ring_offsets[r + 1] - v0;
        for (int64_t i = a; i < b; ++i) {
          const int64_t ind1 = (i == a) ? b - 1 : i - 1, ind2 = i;
          double dy1 = py[ind1], dy2 = py[ind2];
          if ((dy1 < dy && dy2 < dy) || (dy1 > dy && dy2 > dy)) continue;
          double dx1, dx2;
          if (dy1 < dy2) {
            dx1 = px[ind1]; dx2 = px[ind2];
          } else if (dy1 > dy2) {
            double t = dy1; dy1 = dy2; dy2 = t;
            dx2 = px[ind1]; dx1 = px[ind2];
          } else {
            if (px[ind1] > px[ind2]) {
              const int hx1 = (int)floor(px[ind2] + 0.5), hx2 = (int)floor(px[ind1] + 0.5);
              if (hx1 > maxx || hx2 <= minx) continue;
              int x0 = hx1 < minx ? minx : hx1, x1 = hx2 - 1 > maxx ? maxx : hx2 - 1;
              if (x0 <= x1) emit(ctx, p, y, x0, x1);
            }
            continue;
          }
          if (dy < dy2 && dy >= dy1) {
            const double intersect = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1;
            ints[count++] = (int)floor(intersect + 0.5);
          }
        }
      }
      qsort(ints, count, sizeof(int), cmp_int);
      for (int i = 0; i + 1 < count; i += 2) {
        if (ints[i] <= maxx && ints[i + 1] > minx) {
          int x0 = ints[i] < minx ? minx : ints[i];
          int x1 = ints[i + 1] - 1 > maxx ? maxx : ints[i + 1] - 1;
          if (x0 <= x1) emit(ctx, p, y, x0, x1);
        }
      }
    }
    free(px); free(py); free(ints);
  }
}

struct label_ctx { int32_t* labels; int width; };

static void burn_label(void* c, int64_t p, int y, int x0, int x1) {
  struct label_ctx* ctx = (struct label_ctx*)c;
  int32_t* row = ctx->labels + (int64_t)y * ctx->width;
  for (int x = x0; x <= x1; ++x) row[x] = (int32_t)p;
}

/* labels[h*w] must be pre-filled with the "unlabelled" value; polygon p burns index p */
void gm_oracle_burn_index(const double* xy, const int64_t* ring_offsets, const int64_t* poly_offsets,
                          int64_t n_polygons, const double* gt, int height, int width,
                          int32_t* labels) {
  struct label_ctx ctx = {labels, width};
  fill_polygons(xy, ring_offsets, poly_offsets, n_polygons, gt, height, width, burn_label, &ctx);
}

struct span_ctx { int64_t* spans; int64_t capacity; int64_t count; };

static void collect_span(void* c, int64_t p, int y, int x0, int x1) {
  struct span_ctx* ctx = (struct span_ctx*)c;
  if (ctx->count < ctx->capacity) {
    int64_t* s = ctx->spans + 4 * ctx->count;
    s[0] = p; s[1] = y; s[2] = x0; s[3] = x1;
  }
  ctx->count++;
}

/* every (polygon, row, x0, x1) span; returns the number of spans (may exceed capacity) */
int64_t gm_oracle_spans(const double* xy, const int64_t* ring_offsets, const int64_t* poly_offsets,
                        int64_t n_polygons, const double* gt, int height, int width,
                        int64_t* spans, int64_t capacity) {
  struct span_ctx ctx = {spans, capacity, 0};
  fill_polygons(xy, ring_offsets, poly_offsets, n_polygons, gt, height, width, collect_span, &ctx);
  return ctx.count;
}
