// Spatial stencils of raster/spatial.py: HillShade, MovingMax, Dilate, Smooth.
// The source array carries the halo the reference block requested from its
// store (raster/spatial.py:27-108); every kernel reads (t, H+2m, W+2m) and
// writes the cropped (t, H, W).
#include "gm_common.cuh"
#include <cfloat>
#include <cmath>
#include <limits>

namespace gm {

template <typename T> struct Lowest { static __host__ __device__ T value() { return std::numeric_limits<T>::lowest(); } };

// ---------------------------------------------------------------------------------
// HillShade (raster/spatial.py:353-417)
// ---------------------------------------------------------------------------------
// Horn gradient evaluated left to right in the array's dtype, divided by the
// resolution into float32, then the shading in float32 exactly in the order of
// the reference expression; the result is truncated to uint8.
template <typename T> struct HillArith;
template <> struct HillArith<float> {
  typedef float acc;
  static __device__ __forceinline__ float div(float v, double res) { return v / (float)res; }
};
template <> struct HillArith<double> {
  typedef double acc;
  static __device__ __forceinline__ float div(double v, double res) { return (float)(v / res); }
};
template <typename T> struct HillArith {  // integer rasters: arithmetic wraps in T, division in double
  typedef T acc;
  static __device__ __forceinline__ float div(T v, double res) { return (float)((double)v / res); }
};

template <typename T>
__global__ void __launch_bounds__(256)
hillshade_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, T nodata, int has_nodata,
                 T fill, int bands, int H, int W, double xres, double yres,
                 float sin_alt, float cos_alt_zsf, float az, float square_zsf) {
  typedef typename HillArith<T>::acc A;
  const int SW = W + 2;
  const int64_t in_plane = (int64_t)(H + 2) * SW, out_plane = (int64_t)H * W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  for (int b = blockIdx.z; b < bands; b += gridDim.z) {
    const T* p = src + (int64_t)b * in_plane + (int64_t)y * SW + x;
    A s[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        T v = __ldg(p + r * SW + c);
        if (has_nodata && v == nodata) v = fill;
        s[r * 3 + c] = (A)v;
      }
    const A two = (A)2;
    const A gy = ((((s[0] + two * s[1]) + s[2]) - s[6]) - two * s[7]) - s[8];
    const A gx = ((((s[0] + two * s[3]) + s[6]) - s[2]) - two * s[5]) - s[8];
    const float fy = HillArith<T>::div((T)gy, yres), fx = HillArith<T>::div((T)gx, xres);
    const float xx_plus_yy = fx * fx + fy * fy;
    const float aspect = atan2f(fy, fx);
    const float num = sin_alt - (cos_alt_zsf * sqrtf(xx_plus_yy)) * sinf(aspect - az);
    const float cang = num / sqrtf(1.0f + square_zsf * xx_plus_yy);
    uint8_t out = 0;
    if (!(cang <= 0.0f)) out = (uint8_t)(int)(255.0f * cang);
    dst[(int64_t)b * out_plane + (int64_t)y * W + x] = out;
  }
}

// ---------------------------------------------------------------------------------
// MovingMax (raster/spatial.py:192-213): circular footprint, chord per row
// ---------------------------------------------------------------------------------
constexpr int MM_TX = 32, MM_TY = 8, MM_MAXR = 32;

template <typename T>
__global__ void __launch_bounds__(MM_TX * MM_TY)
moving_max_kernel(const T* __restrict__ src, T* __restrict__ dst, T nodata, int has_nodata,
                  int bands, int H, int W, int r, const int* __restrict__ half_width) {
  // shared tile with halo; no data is replaced by the dtype minimum on load
  extern __shared__ __align__(16) unsigned char mm_smem[];
  T* tile = reinterpret_cast<T*>(mm_smem);
  __shared__ int hw[2 * MM_MAXR + 1];
  const int tw = MM_TX + 2 * r, th = MM_TY + 2 * r;
  const int SW = W + 2 * r;
  const int64_t in_plane = (int64_t)(H + 2 * r) * SW, out_plane = (int64_t)H * W;
  const int tid = threadIdx.y * MM_TX + threadIdx.x;
  if (tid < 2 * r + 1) hw[tid] = half_width[tid];
  const int x0 = blockIdx.x * MM_TX, y0 = blockIdx.y * MM_TY;
  const T lowest = Lowest<T>::value();
  for (int b = blockIdx.z; b < bands; b += gridDim.z) {
    __syncthreads();
    for (int i = tid; i < tw * th; i += MM_TX * MM_TY) {
      const int ty = i / tw, tx = i - ty * tw;
      const int gy = y0 + ty, gx = x0 + tx;  // coordinates in the haloed source
      T v = lowest;
      if (gy < H + 2 * r && gx < SW) {
        v = __ldg(src + (int64_t)b * in_plane + (int64_t)gy * SW + gx);
        if (has_nodata && v == nodata) v = lowest;
      }
      tile[i] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
      T best = lowest;
      for (int dy = -r; dy <= r; ++dy) {
        const int w = hw[dy + r];
        const T* row = tile + (threadIdx.y + r + dy) * tw + threadIdx.x + r;
        for (int dx = -w; dx <= w; ++dx) {
          const T v = row[dx];
          best = v > best ? v : best;
        }
      }
      // restore no data only where the centre was no data and nothing was found
      const T centre = __ldg(src + (int64_t)b * in_plane + (int64_t)(y + r) * SW + (x + r));
      if (has_nodata && best == lowest && centre == nodata) best = nodata;
      dst[(int64_t)b * out_plane + (int64_t)y * W + x] = best;
    }
  }
}

// ---------------------------------------------------------------------------------
// Dilate (raster/spatial.py:146-155): 6-connected 3-D cross, values in order
// ---------------------------------------------------------------------------------
constexpr int DILATE_MAX_VALUES = 64;
template <typename T> struct DilateValues { T v[DILATE_MAX_VALUES]; int n; };

template <typename T>
__global__ void __launch_bounds__(256)
dilate_kernel(const T* __restrict__ src, T* __restrict__ dst, int bands, int H, int W,
              const __grid_constant__ DilateValues<T> values) {
  const int SW = W + 2, SH = H + 2;
  const int64_t in_plane = (int64_t)SH * SW, out_plane = (int64_t)H * W;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  for (int b = blockIdx.z; b < bands; b += gridDim.z) {
    const T* c = src + (int64_t)b * in_plane + (int64_t)(y + 1) * SW + (x + 1);
    T nb[7];
    bool ok[7];
    nb[0] = c[0]; ok[0] = true;
    nb[1] = c[-1]; ok[1] = true;      // the halo is part of the array: always in bounds
    nb[2] = c[1]; ok[2] = true;
    nb[3] = c[-SW]; ok[3] = true;
    nb[4] = c[SW]; ok[4] = true;
    ok[5] = b > 0; nb[5] = ok[5] ? c[-in_plane] : c[0];
    ok[6] = b + 1 < bands; nb[6] = ok[6] ? c[in_plane] : c[0];
    int best = -1;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      if (!ok[k]) continue;
      for (int i = values.n - 1; i > best; --i)
        if (nb[k] == values.v[i]) { best = i; break; }
    }
    dst[(int64_t)b * out_plane + (int64_t)y * W + x] = best >= 0 ? values.v[best] : nb[0];
  }
}

// ---------------------------------------------------------------------------------
// Smooth (raster/spatial.py:273-307): scipy.ndimage.gaussian_filter restated
// ---------------------------------------------------------------------------------
// NI_Correlate1D, symmetric branch: tmp = c*w[0] + sum_{k=l..1} (a[-k] + a[+k])*w[k]
// accumulated in double, stored in the array dtype after EACH pass (y, then x).
// mode="constant": taps outside the array read `fill`.
constexpr int SM_TX = 32, SM_TY = 16, SM_MAXL = 64;

template <typename T> __device__ __forceinline__ T from_double(double v) { return (T)v; }

struct SmoothWeights { double wy[SM_MAXL + 1]; double wx[SM_MAXL + 1]; };

template <typename T>
__global__ void __launch_bounds__(SM_TX * SM_TY)
smooth_kernel(const T* __restrict__ src, T* __restrict__ dst, T nodata, int has_nodata, T fill,
              int bands, int SH, int SW, int H, int W, int my, int mx, int ly, int lx,
              const __grid_constant__ SmoothWeights wts) {
  extern __shared__ __align__(16) unsigned char sm_smem[];
  const int tw = SM_TX + 2 * lx, th = SM_TY + 2 * ly;
  T* tile = reinterpret_cast<T*>(sm_smem);     // th x tw input (no data -> fill)
  T* mid = tile + (size_t)th * tw;             // SM_TY x tw result of the y pass
  const int tid = threadIdx.y * SM_TX + threadIdx.x;
  const int64_t in_plane = (int64_t)SH * SW, out_plane = (int64_t)H * W;
  const int x0 = blockIdx.x * SM_TX, y0 = blockIdx.y * SM_TY;  // output coordinates
  for (int b = blockIdx.z; b < bands; b += gridDim.z) {
    __syncthreads();
    for (int i = tid; i < tw * th; i += SM_TX * SM_TY) {
      const int ty = i / tw, tx = i - ty * tw;
      const int gy = y0 + my - ly + ty, gx = x0 + mx - lx + tx;  // source coordinates
      T v = fill;
      if (gy >= 0 && gy < SH && gx >= 0 && gx < SW) {
        v = __ldg(src + (int64_t)b * in_plane + (int64_t)gy * SW + gx);
        if (has_nodata && v == nodata) v = fill;
      }
      tile[i] = v;
    }
    __syncthreads();
    // y pass for every column of the tile (including the x halo)
    for (int i = tid; i < SM_TY * tw; i += SM_TX * SM_TY) {
      const int ty = i / tw, tx = i - ty * tw;
      const T* c = tile + (ty + ly) * tw + tx;
      // rows of the haloed tile that lie outside the source array are `fill`
      // already; rows outside the y-filtered range cannot be requested
      double tmp = (double)c[0] * wts.wy[0];
      for (int k = ly; k >= 1; --k) tmp += ((double)c[-k * tw] + (double)c[k * tw]) * wts.wy[k];
      T r = ly > 0 ? from_double<T>(tmp) : c[0];
      // columns outside the source are padding for the x pass: constant `fill`
      const int gx = x0 + mx - lx + tx;
      if (gx < 0 || gx >= SW) r = fill;
      mid[i] = r;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
      const T* c = mid + threadIdx.y * tw + threadIdx.x + lx;
      T out;
      if (lx > 0) {
        double tmp = (double)c[0] * wts.wx[0];
        for (int k = lx; k >= 1; --k) tmp += ((double)c[-k] + (double)c[k]) * wts.wx[k];
        out = from_double<T>(tmp);
      } else {
        out = c[0];
      }
      dst[(int64_t)b * out_plane + (int64_t)y * W + x] = out;
    }
  }
}

// zoom-back of Smooth's "zoom" mode: ndimage.affine_transform(order=0,
// matrix=diag(1, zy, zx), offset=(0, oy, ox)) -> out[i, j] = in[floor(oy + i*zy + .5), ...]
template <typename T>
__global__ void zoom_nn_kernel(const T* __restrict__ src, T* __restrict__ dst, int bands, int H,
                               int W, double zy, double zx, double oy, double ox) {
  const int64_t plane = (int64_t)H * W, total = plane * bands;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / plane);
    const int64_t r = idx - (int64_t)b * plane;
    const int i = (int)(r / W), j = (int)(r - (int64_t)i * W);
    const double cy = oy + (double)i * zy, cx = ox + (double)j * zx;
    T v = (T)0;  // cval
    if (cy >= 0.0 && cy <= (double)(H - 1) && cx >= 0.0 && cx <= (double)(W - 1)) {
      const int sy = (int)floor(cy + 0.5), sx = (int)floor(cx + 0.5);
      v = src[(int64_t)b * plane + (int64_t)sy * W + sx];
    }
    dst[idx] = v;
  }
}

// ---- host side ---------------------------------------------------------------------
template <typename T> static T read_scalar(const void* p) { T v; memcpy(&v, p, sizeof(T)); return v; }
template <typename T> static T cast_fill(double fill) { return (T)fill; }

static dim3 grid3(int W, int H, int bands, int bx, int by) {
  int gz = bands < 65535 ? bands : 65535;
  return dim3((W + bx - 1) / bx, (H + by - 1) / by, gz);
}

template <typename T>
static int run_hillshade(const Staged& in, Staged& out, const void* nodata, int has_nodata,
                         double fill, int bands, int H, int W, double xres, double yres,
                         double alt_deg, double az_deg, cudaStream_t s) {
  const double alt = alt_deg * (M_PI / 180.0), az = az_deg * (M_PI / 180.0);
  // math.radians(x) = x * (pi / 180) in CPython
  const double zsf = 1.0 / 8.0;
  dim3 block(32, 8);
  hillshade_kernel<T><<<grid3(W, H, bands, 32, 8), block, 0, s>>>(
      (const T*)in.dev, (uint8_t*)out.dev, has_nodata ? read_scalar<T>(nodata) : T(0), has_nodata,
      cast_fill<T>(fill), bands, H, W, xres, yres, (float)sin(alt), (float)(cos(alt) * zsf),
      (float)az, (float)(zsf * zsf));
  GM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int run_moving_max(const Staged& in, Staged& out, const void* nodata, int has_nodata,
                          int bands, int H, int W, int size, cudaStream_t s) {
  const int r = size / 2;
  if (r > MM_MAXR) return fail("gm_moving_max: size too large (max 65)");
  // chord half-widths of the disc x^2 + y^2 < (size/2)^2 (utils.py:536-547)
  std::vector<int> hw(2 * r + 1);
  const double rad2 = (size / 2.0) * (size / 2.0);
  for (int dy = -r; dy <= r; ++dy) {
    int w = -1;
    for (int dx = 0; dx <= r; ++dx)
      if ((double)(dx * dx + dy * dy) < rad2) w = dx;
    hw[dy + r] = w;  // -1: the row is not part of the footprint
  }
  void* dev_hw = nullptr;
  if (upload(&dev_hw, hw.data(), (int64_t)hw.size() * sizeof(int), s)) return 1;
  const size_t smem = (size_t)(MM_TX + 2 * r) * (MM_TY + 2 * r) * sizeof(T);
  auto kernel = moving_max_kernel<T>;
  if (smem > 48 * 1024)
    GM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid3(W, H, bands, MM_TX, MM_TY), dim3(MM_TX, MM_TY), smem, s>>>(
      (const T*)in.dev, (T*)out.dev, has_nodata ? read_scalar<T>(nodata) : T(0), has_nodata, bands,
      H, W, r, (const int*)dev_hw);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(dev_hw, s);
  if (e != cudaSuccess) return fail(std::string("moving_max launch: ") + cudaGetErrorString(e));
  count_launch();
  return 0;
}

template <typename T>
static int run_dilate(const Staged& in, Staged& out, const void* values, int n_values, int bands,
                      int H, int W, cudaStream_t s) {
  if (n_values > DILATE_MAX_VALUES) return fail("gm_dilate: more than 64 values");
  DilateValues<T> dv;
  memset(&dv, 0, sizeof(dv));
  dv.n = n_values;
  memcpy(dv.v, values, sizeof(T) * n_values);
  dilate_kernel<T><<<grid3(W, H, bands, 32, 8), dim3(32, 8), 0, s>>>(
      (const T*)in.dev, (T*)out.dev, bands, H, W, dv);
  GM_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int run_smooth(const Staged& in, Staged& out, const void* nodata, int has_nodata,
                      double fill, const double* wy, int ly, const double* wx, int lx, int bands,
                      int SH, int SW, int H, int W, int my, int mx, int zoom, double zy, double zx,
                      double oy, double ox, cudaStream_t s) {
  if (ly > SM_MAXL || lx > SM_MAXL) return fail("gm_smooth: kernel radius above 64 taps");
  SmoothWeights wts;
  memset(&wts, 0, sizeof(wts));
  // scipy passes weights[::-1]; index k here is the distance from the centre
  for (int k = 0; k <= ly; ++k) wts.wy[k] = wy[ly + k];
  for (int k = 0; k <= lx; ++k) wts.wx[k] = wx[lx + k];
  const size_t smem = ((size_t)(SM_TX + 2 * lx) * (SM_TY + 2 * ly) +
                       (size_t)SM_TY * (SM_TX + 2 * lx)) * sizeof(T);
  auto kernel = smooth_kernel<T>;
  if (smem > 48 * 1024)
    GM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  T* target = (T*)out.dev;
  void* tmp = nullptr;
  if (zoom) {
    GM_CUDA(cudaMallocAsync(&tmp, (size_t)bands * H * W * sizeof(T), s));
    target = (T*)tmp;
  }
  kernel<<<grid3(W, H, bands, SM_TX, SM_TY), dim3(SM_TX, SM_TY), smem, s>>>(
      (const T*)in.dev, target, has_nodata ? read_scalar<T>(nodata) : T(0), has_nodata,
      cast_fill<T>(fill), bands, SH, SW, H, W, my, mx, ly, lx, wts);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (tmp) cudaFreeAsync(tmp, s);
    return fail(std::string("smooth launch: ") + cudaGetErrorString(e));
  }
  count_launch();
  if (zoom) {
    const int64_t total = (int64_t)bands * H * W;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    zoom_nn_kernel<T><<<(unsigned)blocks, 256, 0, s>>>((const T*)tmp, (T*)out.dev, bands, H, W, zy,
                                                        zx, oy, ox);
    e = cudaGetLastError();
    cudaFreeAsync(tmp, s);
    if (e != cudaSuccess) return fail(std::string("zoom launch: ") + cudaGetErrorString(e));
    count_launch();
  }
  return 0;
}

// open both arrays, run `body`, download and release
template <typename F>
static int with_staged(const GmArray* src, GmArray* dst, void* stream, F body) {
  if (ensure_init()) return 1;
  if (!src || !dst) return fail("null array");
  cudaStream_t s = resolve_stream(stream);
  Staged in, out;
  int rc = in.open_input(*src, s);
  if (!rc) rc = out.open_output(*dst, s);
  if (!rc && array_count(*dst) > 0) rc = body(in, out, s);
  if (!rc) rc = out.finish_output();
  const bool sync = out.owned;
  in.release();
  out.release();
  if (!rc && sync) GM_CUDA(cudaStreamSynchronize(s));
  return rc;
}

#define GM_DISPATCH_NUMERIC(dtype, CALL)                                       \
  switch (dtype) {                                                             \
    case GM_U8: case GM_BOOL: { typedef uint8_t T; return CALL; }              \
    case GM_I8:  { typedef int8_t T; return CALL; }                            \
    case GM_U16: { typedef uint16_t T; return CALL; }                          \
    case GM_I16: { typedef int16_t T; return CALL; }                           \
    case GM_U32: { typedef uint32_t T; return CALL; }                          \
    case GM_I32: { typedef int32_t T; return CALL; }                           \
    case GM_I64: { typedef int64_t T; return CALL; }                           \
    case GM_F32: { typedef float T; return CALL; }                             \
    case GM_F64: { typedef double T; return CALL; }                            \
    default: return fail("unsupported dtype");                                 \
  }

}  // namespace gm

using namespace gm;

extern "C" int gm_hillshade(const GmArray* src, GmArray* dst, const void* nodata, int has_nodata,
                            double fill, double xres, double yres, double altitude_deg,
                            double azimuth_deg, void* stream) {
  return with_staged(src, dst, stream, [&](Staged& in, Staged& out, cudaStream_t s) -> int {
    const int bands = (int)dst->shape[0], H = (int)dst->shape[1], W = (int)dst->shape[2];
    if (dst->dtype != GM_U8) return fail("gm_hillshade: output must be uint8");
    if (src->shape[0] != bands || src->shape[1] != H + 2 || src->shape[2] != W + 2)
      return fail("gm_hillshade: source must carry a 1 pixel halo");
    GM_DISPATCH_NUMERIC(src->dtype, run_hillshade<T>(in, out, nodata, has_nodata, fill, bands, H, W,
                                                     xres, yres, altitude_deg, azimuth_deg, s));
  });
}

extern "C" int gm_moving_max(const GmArray* src, GmArray* dst, const void* nodata, int has_nodata,
                             int size, void* stream) {
  return with_staged(src, dst, stream, [&](Staged& in, Staged& out, cudaStream_t s) -> int {
    const int bands = (int)dst->shape[0], H = (int)dst->shape[1], W = (int)dst->shape[2];
    const int r = size / 2;
    if (src->dtype != dst->dtype) return fail("gm_moving_max: dtype mismatch");
    if (src->shape[0] != bands || src->shape[1] != H + 2 * r || src->shape[2] != W + 2 * r)
      return fail("gm_moving_max: source must carry a size//2 pixel halo");
    GM_DISPATCH_NUMERIC(src->dtype, run_moving_max<T>(in, out, nodata, has_nodata, bands, H, W, size, s));
  });
}

extern "C" int gm_dilate(const GmArray* src, GmArray* dst, const void* values, int n_values,
                         void* stream) {
  return with_staged(src, dst, stream, [&](Staged& in, Staged& out, cudaStream_t s) -> int {
    const int bands = (int)dst->shape[0], H = (int)dst->shape[1], W = (int)dst->shape[2];
    if (src->dtype != dst->dtype) return fail("gm_dilate: dtype mismatch");
    if (src->shape[0] != bands || src->shape[1] != H + 2 || src->shape[2] != W + 2)
      return fail("gm_dilate: source must carry a 1 pixel halo");
    GM_DISPATCH_NUMERIC(src->dtype, run_dilate<T>(in, out, values, n_values, bands, H, W, s));
  });
}

extern "C" int gm_smooth(const GmArray* src, GmArray* dst, const void* nodata, int has_nodata,
                         double fill, const double* wy, int ly, const double* wx, int lx, int my,
                         int mx, int zoom, double zy, double zx, double oy, double ox, void* stream) {
  return with_staged(src, dst, stream, [&](Staged& in, Staged& out, cudaStream_t s) -> int {
    const int bands = (int)dst->shape[0], H = (int)dst->shape[1], W = (int)dst->shape[2];
    const int SH = (int)src->shape[1], SW = (int)src->shape[2];
    if (src->dtype != dst->dtype) return fail("gm_smooth: dtype mismatch");
    if (src->shape[0] != bands || SH != H + 2 * my || SW != W + 2 * mx)
      return fail("gm_smooth: source shape must be the output shape plus the margins");
    if (zoom && (my != 0 || mx != 0)) return fail("gm_smooth: zoom mode takes no margins");
    GM_DISPATCH_NUMERIC(src->dtype, run_smooth<T>(in, out, nodata, has_nodata, fill, wy, ly, wx, lx,
                                                  bands, SH, SW, H, W, my, mx, zoom, zy, zx, oy, ox, s));
  });
}
