// Temporary: entry points that are declared in geokernels.h but not built yet.
#include "gm_common.cuh"
using namespace gm;
extern "C" {
int gm_hillshade(const GmArray*, GmArray*, const void*, int, double, double, double, double, double, void*) { return fail("gm_hillshade: not implemented"); }
int gm_moving_max(const GmArray*, GmArray*, const void*, int, int, void*) { return fail("gm_moving_max: not implemented"); }
int gm_dilate(const GmArray*, GmArray*, const void*, int, void*) { return fail("gm_dilate: not implemented"); }
int gm_smooth(const GmArray*, GmArray*, const void*, int, double, const double*, int, const double*, int, int, int, int, double, double, double, double, void*) { return fail("gm_smooth: not implemented"); }
int gm_temporal_aggregate(const GmArray*, GmArray*, const void*, int, int, double, const int32_t*, const int32_t*, int, void*) { return fail("gm_temporal_aggregate: not implemented"); }
int gm_temporal_cumulative(const GmArray*, GmArray*, const void*, int, int, const int32_t*, const int32_t*, const int32_t*, int, void*) { return fail("gm_temporal_cumulative: not implemented"); }
int gm_rasterize_polygons(const GmPolygons*, const double*, const void*, const void*, GmArray*, void*) { return fail("gm_rasterize_polygons: not implemented"); }
int gm_zonal_stats(const GmArray*, const void*, int, const GmPolygons*, const double*, int, double, const float*, int64_t, int64_t, float*, int64_t*, GmZonalPartial*, void*) { return fail("gm_zonal_stats: not implemented"); }
}
