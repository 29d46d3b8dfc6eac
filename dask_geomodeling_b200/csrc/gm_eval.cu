// Fused element-wise evaluator: one HBM pass per program.
//
// Every thread owns V = 4*Q pixels ("quads" of 4 consecutive pixels so that all
// global accesses are 32/64/128-bit vectors, fully coalesced across the warp)
// and interprets the program once for all of them: the instruction stream
// lives in the kernel parameter (constant) bank, decode is warp-uniform, the
// accumulator, the operand and the GM_NREG registers live in the register file.
// Two machines are instantiated: 32-bit slots (classes I32/F32, V = 8) and
// 64-bit slots (all four classes, V = 4).
//
// Reference semantics restated per op: see include/geokernels.h (GmOp) and
// SURVEY.md Appendix B; reference code raster/elemwise.py:235-299, :551-638,
// :726-757 and raster/misc.py:98-123, :208-222, :245-251, :309-328, :387-399,
// :482-515.
#include "gm_common.cuh"
#include <type_traits>

namespace gm {

constexpr int THREADS = 256;

struct DevTable {
  const int64_t* keys;   // sorted keys (int64 or float64 bits)
  const uint64_t* vals;  // 8-byte values
  const uint8_t* hit;    // per entry: 1 mapped, 2 mapped onto the fill value
  int64_t base;
  int32_t n;
  int32_t kind;
};

struct EvalParams {
  int32_t n_instr;
  int32_t n_tables;
  int32_t tables_in_smem;
  int32_t pad;
  int64_t n;
  const void* in[GM_MAX_INPUTS];
  void* out[GM_MAX_OUTPUTS];
  int32_t in_dtype[GM_MAX_INPUTS];
  int32_t out_dtype[GM_MAX_OUTPUTS];
  DevTable tab[GM_MAX_TABLES];
  GmInstr instr[GM_MAX_INSTR];
};

template <int W> struct SlotOf;
template <> struct SlotOf<4> { typedef uint32_t type; };
template <> struct SlotOf<8> { typedef uint64_t type; };

// ---- raw bits <-> typed value -------------------------------------------------
template <typename T> struct Raw;
template <> struct Raw<int32_t> {
  static constexpr int cls = GM_C_I32;
  static __device__ __forceinline__ int32_t get(uint64_t b) { return (int32_t)(uint32_t)b; }
  static __device__ __forceinline__ uint64_t put(int32_t v) { return (uint64_t)(uint32_t)v; }
};
template <> struct Raw<float> {
  static constexpr int cls = GM_C_F32;
  static __device__ __forceinline__ float get(uint64_t b) { return __uint_as_float((uint32_t)b); }
  static __device__ __forceinline__ uint64_t put(float v) { return (uint64_t)__float_as_uint(v); }
};
template <> struct Raw<int64_t> {
  static constexpr int cls = GM_C_I64;
  static __device__ __forceinline__ int64_t get(uint64_t b) { return (int64_t)b; }
  static __device__ __forceinline__ uint64_t put(int64_t v) { return (uint64_t)v; }
};
template <> struct Raw<double> {
  static constexpr int cls = GM_C_F64;
  static __device__ __forceinline__ double get(uint64_t b) { return __longlong_as_double((long long)b); }
  static __device__ __forceinline__ uint64_t put(double v) { return (uint64_t)__double_as_longlong(v); }
};

template <typename T> struct IsFloat { static constexpr bool value = std::is_floating_point<T>::value; };

// value conversion slot(class) -> T; the class switch is warp-uniform
template <typename T, int V, typename S>
__device__ __forceinline__ void widen_all(const S (&src)[V], int cls, T (&dst)[V]) {
  if constexpr (sizeof(S) == 4) {
    if (cls == GM_C_I32) {
#pragma unroll
      for (int i = 0; i < V; ++i) dst[i] = (T)Raw<int32_t>::get(src[i]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) dst[i] = (T)Raw<float>::get(src[i]);
    }
  } else {
    switch (cls) {
      case GM_C_I32:
#pragma unroll
        for (int i = 0; i < V; ++i) dst[i] = (T)Raw<int32_t>::get(src[i]);
        break;
      case GM_C_F32:
#pragma unroll
        for (int i = 0; i < V; ++i) dst[i] = (T)Raw<float>::get(src[i]);
        break;
      case GM_C_I64:
#pragma unroll
        for (int i = 0; i < V; ++i) dst[i] = (T)Raw<int64_t>::get(src[i]);
        break;
      default:
#pragma unroll
        for (int i = 0; i < V; ++i) dst[i] = (T)Raw<double>::get(src[i]);
        break;
    }
  }
}

template <typename T, int V, typename S>
__device__ __forceinline__ void narrow_all(const T (&src)[V], S (&dst)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) dst[i] = (S)Raw<T>::put(src[i]);
}

// numpy astype between classes, in place
template <int V, typename S>
__device__ __forceinline__ void convert_all(S (&x)[V], int from, int to) {
  if (from == to) return;
  if (to == GM_C_I32) { int32_t t[V]; widen_all<int32_t, V>(x, from, t); narrow_all<int32_t, V>(t, x); return; }
  if (to == GM_C_F32) { float t[V]; widen_all<float, V>(x, from, t); narrow_all<float, V>(t, x); return; }
  if constexpr (sizeof(S) == 8) {
    if (to == GM_C_I64) { int64_t t[V]; widen_all<int64_t, V>(x, from, t); narrow_all<int64_t, V>(t, x); return; }
    double t[V]; widen_all<double, V>(x, from, t); narrow_all<double, V>(t, x);
  }
}

// exact `values == no_data_value` in the operand's own class -> bitmask over V
template <typename T, int V, typename S>
__device__ __forceinline__ unsigned eq_mask_t(const S (&x)[V], uint64_t k) {
  unsigned m = 0;
  const T c = Raw<T>::get(k);
#pragma unroll
  for (int i = 0; i < V; ++i) m |= (unsigned)(Raw<T>::get(x[i]) == c) << i;
  return m;
}
template <int V, typename S>
__device__ __forceinline__ unsigned eq_mask(const S (&x)[V], uint64_t k, int cls) {
  if (cls == GM_C_I32) return eq_mask_t<int32_t, V>(x, k);
  if (cls == GM_C_F32) return eq_mask_t<float, V>(x, k);
  if constexpr (sizeof(S) == 8) {
    if (cls == GM_C_I64) return eq_mask_t<int64_t, V>(x, k);
    return eq_mask_t<double, V>(x, k);
  }
  return 0;
}

// ---- arithmetic helpers ---------------------------------------------------------
__device__ __forceinline__ int32_t w_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int64_t w_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
__device__ __forceinline__ float w_add(float a, float b) { return a + b; }
__device__ __forceinline__ double w_add(double a, double b) { return a + b; }
__device__ __forceinline__ int32_t w_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
__device__ __forceinline__ int64_t w_sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
__device__ __forceinline__ float w_sub(float a, float b) { return a - b; }
__device__ __forceinline__ double w_sub(double a, double b) { return a - b; }
__device__ __forceinline__ int32_t w_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
__device__ __forceinline__ int64_t w_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
__device__ __forceinline__ float w_mul(float a, float b) { return a * b; }
__device__ __forceinline__ double w_mul(double a, double b) { return a * b; }

template <typename T> __device__ __forceinline__ T w_div(T a, T b) {
  if constexpr (IsFloat<T>::value) return a / b; else return b == 0 ? T(0) : a / b;
}
// numpy integer power: square-and-multiply with wrap-around.  A negative
// exponent raises in numpy; here it yields 0 (documented limitation).
template <typename T> __device__ __noinline__ T int_pow(T base, T e) {
  if (e < 0) return 0;
  T r = 1;
  while (e) {
    if (e & 1) r = w_mul(r, base);
    e >>= 1;
    if (e) base = w_mul(base, base);
  }
  return r;
}
// transcendentals are kept out of line: they are rare, large, and would
// otherwise dictate the register allocation of the whole interpreter
__device__ __noinline__ float  t_pow(float a, float b) { return powf(a, b); }
__device__ __noinline__ double t_pow(double a, double b) { return pow(a, b); }
__device__ __noinline__ float  t_exp(float a) { return expf(a); }
__device__ __noinline__ double t_exp(double a) { return exp(a); }
__device__ __noinline__ float  t_log(float a) { return logf(a); }
__device__ __noinline__ double t_log(double a) { return log(a); }
__device__ __noinline__ float  t_log10(float a) { return log10f(a); }
__device__ __noinline__ double t_log10(double a) { return log10(a); }

template <typename T> __device__ __forceinline__ T w_pow(T a, T b) {
  if constexpr (IsFloat<T>::value) return t_pow(a, b); else return int_pow<T>(a, b);
}
template <typename T> __device__ __forceinline__ T w_exp(T a) {
  if constexpr (IsFloat<T>::value) return t_exp(a); else return a;
}
template <typename T> __device__ __forceinline__ T w_log(T a) {
  if constexpr (IsFloat<T>::value) return t_log(a); else return a;
}
template <typename T> __device__ __forceinline__ T w_log10(T a) {
  if constexpr (IsFloat<T>::value) return t_log10(a); else return a;
}
template <typename T> __device__ __forceinline__ bool finite_(T v) {
  if constexpr (IsFloat<T>::value) return isfinite(v); else return true;
}
template <typename T> __device__ __forceinline__ T abs_(T v) {
  if constexpr (IsFloat<T>::value) return fabs(v); else return v < 0 ? -v : v;
}
// np.isclose(x, y): less_equal(abs(x - y), tol) & isfinite(y) | (x == y)
template <typename T> __device__ __forceinline__ bool close_(T x, T y, T tol, bool y_finite) {
  return ((abs_(w_sub(x, y)) <= tol) && y_finite) || (x == y);
}

// ---- typed execution of one instruction over V pixels ----------------------------
// HOMO: acc (and b) already hold class T, so the sentinel tests are plain typed
// compares folded into the op; otherwise operands are widened and the sentinel
// tests were done beforehand in their own classes (inv_a / inv_b bitmasks).
template <typename T, int W, int V, bool HOMO, typename S>
__device__ __forceinline__ void exec_typed(const GmInstr& in, S (&acc)[V], S (&b)[V],
                                           unsigned inv_a, unsigned inv_b,
                                           const DevTable* __restrict__ tabs) {
  constexpr bool kStorable = (sizeof(T) <= W);  // a result of class T fits a slot
  T xa[V], xb[V];
  if constexpr (HOMO) {
#pragma unroll
    for (int i = 0; i < V; ++i) { xa[i] = Raw<T>::get(acc[i]); xb[i] = Raw<T>::get(b[i]); }
  } else {
    widen_all<T, V>(acc, in.cls_a, xa);
    if (in.src_kind != GM_SRC_NONE) widen_all<T, V>(b, in.cls_b, xb);
    else {
#pragma unroll
      for (int i = 0; i < V; ++i) xb[i] = T(0);
    }
  }
  const bool fa = in.flags & GM_F_ND_A, fb = in.flags & GM_F_ND_B;
  const T nda = Raw<T>::get(in.k[1]), ndb = Raw<T>::get(in.k[2]);
  const int op = in.op;

#define GM_INV(i)   (HOMO ? ((fa && xa[i] == nda) || (fb && xb[i] == ndb)) \
                          : (bool)(((inv_a | inv_b) >> (i)) & 1u))
#define GM_MATH(EXPR)                                                          \
  if constexpr (kStorable) {                                                   \
    const T fill = Raw<T>::get(in.k[3]);                                       \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                            \
      T r = (EXPR);                                                            \
      bool bad = GM_INV(i) || !finite_(r);                                     \
      acc[i] = (S)Raw<T>::put(bad ? fill : r);                                 \
    }                                                                          \
  }
#define GM_CMP(EXPR)                                                           \
  {                                                                            \
    const S fill = (S)in.k[3];                                                 \
    _Pragma("unroll") for (int i = 0; i < V; ++i) {                            \
      bool r = (EXPR);                                                         \
      acc[i] = GM_INV(i) ? fill : (S)(r ? 1u : 0u);                            \
    }                                                                          \
  }

  switch (op) {
    case GM_OP_ADD:  GM_MATH(w_add(xa[i], xb[i])) break;
    case GM_OP_SUB:  GM_MATH(w_sub(xa[i], xb[i])) break;
    case GM_OP_RSUB: GM_MATH(w_sub(xb[i], xa[i])) break;
    case GM_OP_MUL:  GM_MATH(w_mul(xa[i], xb[i])) break;
    case GM_OP_DIV:  GM_MATH(w_div(xa[i], xb[i])) break;
    case GM_OP_RDIV: GM_MATH(w_div(xb[i], xa[i])) break;
    case GM_OP_POW:  GM_MATH(w_pow(xa[i], xb[i])) break;
    case GM_OP_RPOW: GM_MATH(w_pow(xb[i], xa[i])) break;
    case GM_OP_EXP:   GM_MATH(w_exp(xa[i])) break;
    case GM_OP_LOG:   GM_MATH(w_log(xa[i])) break;
    case GM_OP_LOG10: GM_MATH(w_log10(xa[i])) break;
    case GM_OP_EQ: GM_CMP(xa[i] == xb[i]) break;
    case GM_OP_NE: GM_CMP(xa[i] != xb[i]) break;
    case GM_OP_GT: GM_CMP(xa[i] > xb[i]) break;
    case GM_OP_GE: GM_CMP(xa[i] >= xb[i]) break;
    case GM_OP_LT: GM_CMP(xa[i] < xb[i]) break;
    case GM_OP_LE: GM_CMP(xa[i] <= xb[i]) break;
    case GM_OP_MASK: {
      // raster/misc.py:208-222: data -> value (k0), no data -> fill (k3)
      const T tol = Raw<T>::get(in.k[4]);
      const bool has = in.flags & GM_F_ND_T, cl = in.flags & GM_F_CLOSE, fin = in.flags & GM_F_ND_FINITE;
      const S val = (S)in.k[0], fill = (S)in.k[3];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        bool nod = has && (cl ? close_(xa[i], nda, tol, fin) : (xa[i] == nda));
        acc[i] = nod ? fill : val;
      }
      break;
    }
    case GM_OP_OVERLAY: {
      // raster/elemwise.py:752-755: values[index] = data[index]
      const T tol = Raw<T>::get(in.k[4]);
      const bool has = in.flags & GM_F_ND_T, cl = in.flags & GM_F_CLOSE, fin = in.flags & GM_F_ND_FINITE;
      unsigned data = 0;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        bool nod = has && (cl ? close_(xb[i], ndb, tol, fin) : (xb[i] == ndb));
        data |= (unsigned)(!nod) << i;
      }
      convert_all<V>(b, in.cls_b, in.cls_out);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = ((data >> i) & 1u) ? b[i] : acc[i];
      break;
    }
    case GM_OP_MASKBELOW: {
      // raster/misc.py:249-250 (k0 threshold in class T, k5 sentinel as slot bits)
      const T thr = Raw<T>::get(in.k[0]);
      const S nd = (S)in.k[5];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = (xa[i] < thr) ? nd : acc[i];
      break;
    }
    case GM_OP_STEP: {
      // raster/misc.py:316-326 (k0 location in class T; k2/k3/k4 left/at/right as slot bits)
      const T loc = Raw<T>::get(in.k[0]);
      const S left = (S)in.k[2], at = (S)in.k[3], right = (S)in.k[4];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        S r = acc[i];
        if (xa[i] < loc) r = left;
        if (xa[i] == loc) r = at;
        if (xa[i] > loc) r = right;
        acc[i] = ((inv_a >> i) & 1u) ? acc[i] : r;
      }
      break;
    }
    case GM_OP_CLASSIFY: {
      // np.digitize(values, bins, right) (raster/misc.py:396), bins ascending
      const DevTable& t = tabs[in.aux];
      const bool right = in.flags & GM_F_RIGHT;
      const S fill = (S)in.k[3];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        int lo = 0, hi = t.n;
        const T x = xa[i];
        if (x != x) lo = t.n;  // NaN sorts last
        else
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            T e = Raw<T>::get((uint64_t)t.keys[mid]);
            bool go = right ? (e < x) : (e <= x);
            if (go) lo = mid + 1; else hi = mid;
          }
        acc[i] = ((inv_a >> i) & 1u) ? fill : (S)(uint32_t)lo;
      }
      break;
    }
    default: break;
  }
#undef GM_MATH
#undef GM_CMP
#undef GM_INV
}

// Reclassify (raster/misc.py:505-514): mapped -> target, else fill (select) or
// astype(values).  With GM_F_ND_T only the "result has data" boolean is produced
// (enough when the result merely masks another raster), which keeps the whole
// program in 32-bit slots.
template <int W, int V, typename S>
__device__ __forceinline__ void exec_reclass(const GmInstr& in, S (&acc)[V],
                                             const DevTable* __restrict__ tabs) {
  const DevTable& t = tabs[in.aux];
  const bool select = in.flags & GM_F_SELECT;
  const bool nd_only = in.flags & GM_F_ND_T;
  int64_t key[V];
  widen_all<int64_t, V>(acc, in.cls_a, key);
  if (!nd_only) convert_all<V>(acc, in.cls_a, in.cls_out);
  const S fill = (S)in.k[3];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    int found = 0;  // 0 miss, 1 mapped, 2 mapped onto the fill value
    uint64_t val = 0;
    if (t.kind == GM_TABLE_DENSE) {
      const int64_t idx = key[i] - t.base;
      if (idx >= 0 && idx < t.n) {
        found = t.hit[idx];
        if (!nd_only && found) val = t.vals[idx];
      }
    } else {
      int lo = 0, hi = t.n;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (t.keys[mid] < key[i]) lo = mid + 1; else hi = mid;
      }
      if (lo < t.n && t.keys[lo] == key[i]) {
        found = t.hit ? t.hit[lo] : 1;
        if (!nd_only) val = t.vals[lo];
      }
    }
    if (nd_only) acc[i] = (S)((found == 1 || (found == 0 && !select)) ? 1u : 0u);
    else if constexpr (W == 8) acc[i] = found ? (S)val : (select ? fill : acc[i]);
  }
}

// ---- global memory access ---------------------------------------------------------
template <typename S>
__device__ __forceinline__ void load_quad(const void* __restrict__ base, int dtype, int64_t pix,
                                          int64_t n, S* dst) {
  // dst[0..3] <- pixels pix..pix+3 as the natural class of `dtype`
  if (pix + 4 <= n) {
    switch (dtype) {
      case GM_BOOL: case GM_U8: {
        uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(base) + (pix >> 2));
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = (S)((w >> (8 * e)) & 0xffu);
        break;
      }
      case GM_I8: {
        uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(base) + (pix >> 2));
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = (S)(uint32_t)(int32_t)(int8_t)((w >> (8 * e)) & 0xffu);
        break;
      }
      case GM_U16: {
        uint2 w = __ldcs(reinterpret_cast<const uint2*>(base) + (pix >> 2));
        dst[0] = (S)(w.x & 0xffffu); dst[1] = (S)(w.x >> 16);
        dst[2] = (S)(w.y & 0xffffu); dst[3] = (S)(w.y >> 16);
        break;
      }
      case GM_I16: {
        uint2 w = __ldcs(reinterpret_cast<const uint2*>(base) + (pix >> 2));
        dst[0] = (S)(uint32_t)(int32_t)(int16_t)(w.x & 0xffffu);
        dst[1] = (S)(uint32_t)(int32_t)(int16_t)(w.x >> 16);
        dst[2] = (S)(uint32_t)(int32_t)(int16_t)(w.y & 0xffffu);
        dst[3] = (S)(uint32_t)(int32_t)(int16_t)(w.y >> 16);
        break;
      }
      case GM_I32: case GM_F32: case GM_U32: {  // U32: natural class I64, zero extended
        uint4 w = __ldcs(reinterpret_cast<const uint4*>(base) + (pix >> 2));
        dst[0] = (S)w.x; dst[1] = (S)w.y; dst[2] = (S)w.z; dst[3] = (S)w.w;
        break;
      }
      default: {  // I64 / F64
        if constexpr (sizeof(S) == 8) {
          const ulonglong2* p = reinterpret_cast<const ulonglong2*>(base) + (pix >> 1);
          ulonglong2 a = __ldcs(p), c = __ldcs(p + 1);
          dst[0] = a.x; dst[1] = a.y; dst[2] = c.x; dst[3] = c.y;
        }
        break;
      }
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int64_t q = pix + e;
    S v = 0;
    if (q < n) {
      switch (dtype) {
        case GM_BOOL: case GM_U8: v = (S)reinterpret_cast<const uint8_t*>(base)[q]; break;
        case GM_I8:  v = (S)(uint32_t)(int32_t)reinterpret_cast<const int8_t*>(base)[q]; break;
        case GM_U16: v = (S)reinterpret_cast<const uint16_t*>(base)[q]; break;
        case GM_I16: v = (S)(uint32_t)(int32_t)reinterpret_cast<const int16_t*>(base)[q]; break;
        case GM_I32: case GM_F32: case GM_U32: v = (S)reinterpret_cast<const uint32_t*>(base)[q]; break;
        default:
          if constexpr (sizeof(S) == 8) v = reinterpret_cast<const uint64_t*>(base)[q];
          break;
      }
    }
    dst[e] = v;
  }
}

template <typename S>
__device__ __forceinline__ void store_quad(void* __restrict__ base, int dtype, int64_t pix,
                                           int64_t n, const S* src) {
  if (pix + 4 <= n) {
    switch (dtype) {
      case GM_BOOL: case GM_U8: case GM_I8: {
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) w |= ((uint32_t)src[e] & 0xffu) << (8 * e);
        __stcs(reinterpret_cast<uint32_t*>(base) + (pix >> 2), w);
        break;
      }
      case GM_U16: case GM_I16: {
        uint2 w;
        w.x = ((uint32_t)src[0] & 0xffffu) | (((uint32_t)src[1] & 0xffffu) << 16);
        w.y = ((uint32_t)src[2] & 0xffffu) | (((uint32_t)src[3] & 0xffffu) << 16);
        __stcs(reinterpret_cast<uint2*>(base) + (pix >> 2), w);
        break;
      }
      case GM_I32: case GM_F32: case GM_U32: {
        uint4 w = make_uint4((uint32_t)src[0], (uint32_t)src[1], (uint32_t)src[2], (uint32_t)src[3]);
        __stcs(reinterpret_cast<uint4*>(base) + (pix >> 2), w);
        break;
      }
      default: {
        if constexpr (sizeof(S) == 8) {
          ulonglong2* p = reinterpret_cast<ulonglong2*>(base) + (pix >> 1);
          __stcs(p, make_ulonglong2(src[0], src[1]));
          __stcs(p + 1, make_ulonglong2(src[2], src[3]));
        }
        break;
      }
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int64_t q = pix + e;
    if (q < n) {
      switch (dtype) {
        case GM_BOOL: case GM_U8: case GM_I8: reinterpret_cast<uint8_t*>(base)[q] = (uint8_t)src[e]; break;
        case GM_U16: case GM_I16: reinterpret_cast<uint16_t*>(base)[q] = (uint16_t)src[e]; break;
        case GM_I32: case GM_F32: case GM_U32: reinterpret_cast<uint32_t*>(base)[q] = (uint32_t)src[e]; break;
        default:
          if constexpr (sizeof(S) == 8) reinterpret_cast<uint64_t*>(base)[q] = src[e];
          break;
      }
    }
  }
}

template <int V, typename S> __device__ __forceinline__ void copy_all(S (&d)[V], const S (&s)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) d[i] = s[i];
}

template <typename T, int W, int V, typename S>
__device__ __forceinline__ void dispatch_typed(const GmInstr& in, S (&acc)[V], S (&b)[V],
                                               const DevTable* __restrict__ tabs) {
  if constexpr (sizeof(T) <= W) {
    const bool homo = in.cls_a == Raw<T>::cls &&
                      (in.src_kind == GM_SRC_NONE || in.cls_b == Raw<T>::cls) &&
                      in.op != GM_OP_STEP && in.op != GM_OP_CLASSIFY;
    if (homo) { exec_typed<T, W, V, true, S>(in, acc, b, 0u, 0u, tabs); return; }
  }
  unsigned inv_a = 0, inv_b = 0;
  if (in.flags & GM_F_ND_A) inv_a = eq_mask<V>(acc, in.k[1], in.cls_a);
  if (in.flags & GM_F_ND_B) inv_b = eq_mask<V>(b, in.k[2], in.cls_b);
  exec_typed<T, W, V, false, S>(in, acc, b, inv_a, inv_b, tabs);
}

// ---- the kernel ---------------------------------------------------------------------
template <int W, int Q>
__global__ void __launch_bounds__(THREADS, 2)
eval_kernel(const __grid_constant__ EvalParams p) {
  typedef typename SlotOf<W>::type S;
  constexpr int V = 4 * Q;
  constexpr int TILE = THREADS * V;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ DevTable tabs[GM_MAX_TABLES];

  // stage the lookup tables in shared memory
  if (threadIdx.x == 0) {
    size_t off = 0;
    for (int t = 0; t < p.n_tables; ++t) {
      DevTable d = p.tab[t];
      if (p.tables_in_smem) {
        if (d.keys) { d.keys = reinterpret_cast<const int64_t*>(smem + off); off += (size_t)d.n * 8; }
        if (d.vals) { d.vals = reinterpret_cast<const uint64_t*>(smem + off); off += (size_t)d.n * 8; }
        if (d.hit)  { d.hit = smem + off; off += ((size_t)d.n + 15) / 16 * 16; }
      }
      tabs[t] = d;
    }
  }
  __syncthreads();
  if (p.tables_in_smem) {
    for (int t = 0; t < p.n_tables; ++t) {
      const DevTable g = p.tab[t];
      const DevTable s = tabs[t];
      if (g.keys) for (int i = threadIdx.x; i < g.n; i += THREADS) const_cast<int64_t*>(s.keys)[i] = g.keys[i];
      if (g.vals) for (int i = threadIdx.x; i < g.n; i += THREADS) const_cast<uint64_t*>(s.vals)[i] = g.vals[i];
      if (g.hit)  for (int i = threadIdx.x; i < g.n; i += THREADS) const_cast<uint8_t*>(s.hit)[i] = g.hit[i];
    }
    __syncthreads();
  }

  const int64_t n = p.n;
  const int64_t n_tiles = (n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * TILE + (int64_t)threadIdx.x * 4;
    S acc[V], b[V], r0[V], r1[V], r2[V], r3[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { acc[i] = 0; b[i] = 0; r0[i] = 0; r1[i] = 0; r2[i] = 0; r3[i] = 0; }

    for (int pc = 0; pc < p.n_instr; ++pc) {
      const GmInstr& in = p.instr[pc];
      // 1. materialise operand b
      switch (in.src_kind) {
        case GM_SRC_REG:
          switch (in.src) {
            case 0: copy_all<V>(b, r0); break;
            case 1: copy_all<V>(b, r1); break;
            case 2: copy_all<V>(b, r2); break;
            default: copy_all<V>(b, r3); break;
          }
          break;
        case GM_SRC_INPUT: {
          const void* ptr = p.in[in.src];
          const int dt = p.in_dtype[in.src];
#pragma unroll
          for (int q = 0; q < Q; ++q)
            load_quad<S>(ptr, dt, base + (int64_t)q * THREADS * 4, n, &b[4 * q]);
          break;
        }
        case GM_SRC_IMM: {
          const S c = (S)in.k[0];
#pragma unroll
          for (int i = 0; i < V; ++i) b[i] = c;
          break;
        }
        default: break;
      }
      // 2. untyped instructions
      const int op = in.op;
      if (op == GM_OP_LOAD) {
        convert_all<V>(b, in.cls_b, in.cls_out);
        copy_all<V>(acc, b);
        continue;
      }
      if (op == GM_OP_ST) {
        switch (in.aux) {
          case 0: copy_all<V>(r0, acc); break;
          case 1: copy_all<V>(r1, acc); break;
          case 2: copy_all<V>(r2, acc); break;
          default: copy_all<V>(r3, acc); break;
        }
        continue;
      }
      if (op == GM_OP_OUT) {
        void* ptr = p.out[in.aux];
        const int dt = p.out_dtype[in.aux];
#pragma unroll
        for (int q = 0; q < Q; ++q)
          store_quad<S>(ptr, dt, base + (int64_t)q * THREADS * 4, n, &acc[4 * q]);
        continue;
      }
      if (op == GM_OP_CVT) { convert_all<V>(acc, in.cls_a, in.cls_out); continue; }
      if (op == GM_OP_ISDATA || op == GM_OP_ISNODATA) {
        unsigned m = (in.flags & GM_F_ND_A) ? eq_mask<V>(acc, in.k[1], in.cls_a) : 0u;
        if (op == GM_OP_ISDATA) m = ~m;
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = (S)((m >> i) & 1u);
        continue;
      }
      if (op == GM_OP_CLIP) {
        unsigned masked = 0;
        if (in.flags & GM_F_B_BOOL) {
#pragma unroll
          for (int i = 0; i < V; ++i) masked |= (unsigned)((uint32_t)b[i] == 0u) << i;
        } else if (in.flags & GM_F_ND_B) {
          masked = eq_mask<V>(b, in.k[2], in.cls_b);
        }
        const S nd = (S)in.k[1];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = ((masked >> i) & 1u) ? nd : acc[i];
        continue;
      }
      if (op >= GM_OP_AND && op <= GM_OP_NOT) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const bool x = (uint32_t)acc[i] != 0u, y = (uint32_t)b[i] != 0u;
          bool r = op == GM_OP_AND ? (x && y) : op == GM_OP_OR ? (x || y) : op == GM_OP_XOR ? (x != y) : !x;
          acc[i] = (S)(r ? 1u : 0u);
        }
        continue;
      }
      if (op == GM_OP_RECLASS) { exec_reclass<W, V, S>(in, acc, tabs); continue; }
      // 3. typed instructions
      if (in.cls == GM_C_I32) { dispatch_typed<int32_t, W, V, S>(in, acc, b, tabs); continue; }
      if (in.cls == GM_C_F32) { dispatch_typed<float, W, V, S>(in, acc, b, tabs); continue; }
      if constexpr (W == 8) {
        if (in.cls == GM_C_I64) dispatch_typed<int64_t, W, V, S>(in, acc, b, tabs);
        else dispatch_typed<double, W, V, S>(in, acc, b, tabs);
      }
    }
  }
}

static bool wide(int cls) { return cls == GM_C_I64 || cls == GM_C_F64; }

static int validate(const GmProgram* prog) {
  if (!prog) return fail("gm_eval_program: null program");
  if (prog->n_instr < 1 || prog->n_instr > GM_MAX_INSTR) return fail("gm_eval_program: bad instruction count");
  if (prog->n_inputs < 0 || prog->n_inputs > GM_MAX_INPUTS) return fail("gm_eval_program: too many inputs");
  if (prog->n_outputs < 1 || prog->n_outputs > GM_MAX_OUTPUTS) return fail("gm_eval_program: bad output count");
  if (prog->n_tables < 0 || prog->n_tables > GM_MAX_TABLES) return fail("gm_eval_program: too many tables");
  if (prog->word != 4 && prog->word != 8) return fail("gm_eval_program: word must be 4 or 8");
  for (int i = 0; i < prog->n_instr; ++i) {
    const GmInstr& in = prog->instr[i];
    if (in.op >= GM_OP_COUNT_) return fail("gm_eval_program: unknown opcode");
    if (in.src_kind == GM_SRC_REG && in.src >= GM_NREG) return fail("gm_eval_program: bad register");
    if (in.src_kind == GM_SRC_INPUT && in.src >= prog->n_inputs) return fail("gm_eval_program: bad input index");
    if (in.op == GM_OP_ST && in.aux >= GM_NREG) return fail("gm_eval_program: bad register");
    if (in.op == GM_OP_OUT && (int)in.aux >= prog->n_outputs) return fail("gm_eval_program: bad output index");
    if ((in.op == GM_OP_CLASSIFY || in.op == GM_OP_RECLASS) && (int)in.aux >= prog->n_tables)
      return fail("gm_eval_program: bad table index");
    if (in.cls > GM_C_F64 || in.cls_a > GM_C_F64 || in.cls_b > GM_C_F64 || in.cls_out > GM_C_F64)
      return fail("gm_eval_program: bad class");
    if (prog->word == 4) {
      if (wide(in.cls_a) || wide(in.cls_out) || wide(in.cls) ||
          (in.src_kind != GM_SRC_NONE && wide(in.cls_b)))
        return fail("gm_eval_program: 64-bit class in a 32-bit program");
    }
  }
  return 0;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_eval_program(const GmProgram* prog, const GmArray* inputs, GmArray* outputs,
                               int64_t n_pixels, void* stream) {
  if (ensure_init()) return 1;
  if (validate(prog)) return 1;
  if (n_pixels < 0) return fail("gm_eval_program: negative pixel count");
  cudaStream_t s = resolve_stream(stream);

  EvalParams p;
  memset(&p, 0, sizeof(p));
  p.n_instr = prog->n_instr;
  p.n_tables = prog->n_tables;
  p.n = n_pixels;
  memcpy(p.instr, prog->instr, sizeof(GmInstr) * prog->n_instr);

  Staged sin[GM_MAX_INPUTS], sout[GM_MAX_OUTPUTS];
  std::vector<void*> scratch;
  int rc = 0;
  auto cleanup = [&]() {
    for (int i = 0; i < prog->n_inputs; ++i) sin[i].release();
    for (int i = 0; i < prog->n_outputs; ++i) sout[i].release();
    for (void* d : scratch) cudaFreeAsync(d, s);
  };

  for (int i = 0; i < prog->n_inputs && !rc; ++i) {
    if (array_count(inputs[i]) != n_pixels) { rc = fail("gm_eval_program: input size mismatch"); break; }
    if (((uintptr_t)inputs[i].data & 15u) && inputs[i].space == GM_DEVICE) {
      rc = fail("gm_eval_program: device input not 16-byte aligned"); break;
    }
    rc = sin[i].open_input(inputs[i], s);
    p.in[i] = sin[i].dev;
    p.in_dtype[i] = inputs[i].dtype;
  }
  for (int i = 0; i < prog->n_outputs && !rc; ++i) {
    if (array_count(outputs[i]) != n_pixels) { rc = fail("gm_eval_program: output size mismatch"); break; }
    if (((uintptr_t)outputs[i].data & 15u) && outputs[i].space == GM_DEVICE) {
      rc = fail("gm_eval_program: device output not 16-byte aligned"); break;
    }
    rc = sout[i].open_output(outputs[i], s);
    p.out[i] = sout[i].dev;
    p.out_dtype[i] = outputs[i].dtype;
  }
  size_t table_bytes = 0;
  for (int t = 0; t < prog->n_tables && !rc; ++t) {
    const GmTable& g = prog->tables[t];
    DevTable& d = p.tab[t];
    d.n = g.n; d.kind = g.kind; d.base = g.base;
    void* dev = nullptr;
    if (g.keys) {
      rc = upload(&dev, g.keys, (int64_t)g.n * 8, s); if (rc) break;
      scratch.push_back(dev); d.keys = (const int64_t*)dev; table_bytes += (size_t)g.n * 8;
    }
    if (g.vals) {
      rc = upload(&dev, g.vals, (int64_t)g.n * 8, s); if (rc) break;
      scratch.push_back(dev); d.vals = (const uint64_t*)dev; table_bytes += (size_t)g.n * 8;
    }
    if (g.hit) {
      rc = upload(&dev, g.hit, (int64_t)g.n, s); if (rc) break;
      scratch.push_back(dev); d.hit = (const uint8_t*)dev; table_bytes += ((size_t)g.n + 15) / 16 * 16;
    }
  }
  if (rc) { cleanup(); return 1; }

  if (n_pixels > 0) {
    p.tables_in_smem = (table_bytes > 0 && table_bytes <= 96 * 1024) ? 1 : 0;
    const size_t smem = p.tables_in_smem ? table_bytes : 0;
    const int V = prog->word == 4 ? 8 : 4;
    const int64_t tile = (int64_t)THREADS * V;
    const int64_t n_tiles = (n_pixels + tile - 1) / tile;
    auto kernel = prog->word == 4 ? eval_kernel<4, 2> : eval_kernel<8, 1>;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    if (e == cudaSuccess)
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem);
    if (e != cudaSuccess) { cleanup(); return fail(std::string("eval launch setup: ") + cudaGetErrorString(e)); }
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    kernel<<<(unsigned)grid, THREADS, smem, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { cleanup(); return fail(std::string("eval kernel launch: ") + cudaGetErrorString(e)); }
    count_launch();
  }
  bool staged_out = false;
  for (int i = 0; i < prog->n_outputs; ++i) {
    if (sout[i].finish_output()) { cleanup(); return 1; }
    staged_out |= sout[i].owned;
  }
  cleanup();
  if (staged_out) GM_CUDA(cudaStreamSynchronize(s));
  return 0;
}
