// Temporary: entry points that are declared in geokernels.h but not built yet.
#include "gm_common.cuh"
using namespace gm;
extern "C" {
int gm_rasterize_polygons(const GmPolygons*, const double*, const void*, const void*, GmArray*, void*) { return fail("gm_rasterize_polygons: not implemented"); }
int gm_zonal_stats(const GmArray*, const void*, int, const GmPolygons*, const double*, int, double, const float*, int64_t, int64_t, float*, int64_t*, GmZonalPartial*, void*) { return fail("gm_zonal_stats: not implemented"); }
}
