// osc_warp.cuh -- the tiny SIMT abstraction osc_core3.cuh is written against.
//
// On the GPU a `Var<T>` is one register of the calling thread, `OSC_LANES(l)` runs its body
// once with l = the thread's lane, and the collectives are single warp instructions
// (SHFL, DMMA, VOTE).  On the host (tests/host_core only -- the product never runs there) a
// `Var<T>` holds the value of all 32 lanes, `OSC_LANES(l)` loops l = 0..31, and the
// collectives are spelled out with the same data movement and the same summation order, so
// `pytest -m "not gpu"` checks the very source the kernel compiles.
//
// Rules the core follows so that both readings agree: no cross-lane communication inside an
// OSC_LANES body; values that live across bodies are Vars; control flow outside the bodies
// depends only on warp-uniform values.
#pragma once

#include <math.h>
#include <string.h>

#ifndef OSC_HD
#if defined(__CUDACC__)
#define OSC_HD __host__ __device__ __forceinline__
#else
#define OSC_HD inline
#endif
#endif

// keeps the compiler from moving memory accesses across this point (used where hoisting
// every load of a long unrolled phase to its top costs more registers than it hides latency)
#if defined(__CUDA_ARCH__)
#define OSC_COMPILER_BARRIER() asm volatile("" ::: "memory")
#else
#define OSC_COMPILER_BARRIER() ((void)0)
#endif

namespace osc {

// a value the compiler must keep in a register instead of recomputing it from the lane index
// wherever it is used (index arithmetic inside the iteration loop)
OSC_HD int osc_opaque(int v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(v));
#endif
  return v;
}

#if defined(__CUDA_ARCH__)

template <class T>
struct Var {
  T v;
  OSC_HD T& operator[](int) { return v; }
  OSC_HD const T& operator[](int) const { return v; }
};
#define OSC_LANES(l) if (const int l = lane0; true)

struct Warp {
  static constexpr unsigned kFull = 0xffffffffu;
  static OSC_HD void sync() { __syncwarp(); }
  // dst[l] = src[l ^ 16]
  static OSC_HD void xchg16(Var<double>& dst, const Var<double>& src) {
    dst.v = __shfl_xor_sync(kFull, src.v, 16);
  }
  // dst[l] = src[(l & ~3) | r]   (broadcast inside a group of four lanes)
  static OSC_HD void group4(Var<double>& dst, const Var<double>& src, int r) {
    dst.v = __shfl_sync(kFull, src.v, r, 4);
  }
  // the value of lane `from` in every lane
  static OSC_HD double bcast(const Var<double>& src, int from) {
    return __shfl_sync(kFull, src.v, from);
  }
  static OSC_HD double sum(const Var<double>& a) {
    double v = a.v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
  }
  // Maximum of NON-NEGATIVE values (a NaN never becomes the maximum, like OSQP's
  // vec_norm_inf): for non-negative doubles the numeric order is the unsigned order of the
  // (high word, low word) pair, so two integer warp reductions (REDUX) replace five rounds of
  // 64-bit shuffle + compare + select.
  static OSC_HD double max(const Var<double>& a) {
    const double v = a.v == a.v ? a.v : 0.0;
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned himax = __reduce_max_sync(kFull, hi);
    const unsigned lomax = __reduce_max_sync(kFull, hi == himax ? lo : 0u);
    return __hiloint2double((int)himax, (int)lomax);
  }
  static OSC_HD unsigned ballot(const Var<bool>& p) { return __ballot_sync(kFull, p.v); }
  // Warp-wide maxima of NQ non-negative quantities; every lane receives all of them.  Two
  // integer warp reductions per quantity (see max()); `scratch` is unused on the device (the
  // results arrive in uniform registers), the host emulation keeps the interface.
  template <int NQ>
  static OSC_HD void maxn(Var<double> (&m)[NQ], double* out, double* scratch, int lane) {
    (void)scratch;
    (void)lane;
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = max(m[q]);
  }
  static OSC_HD void max16(Var<double> (&m)[16], double* out, double* scratch, int lane) {
    maxn<16>(m, out, scratch, lane);
  }
  // FP64 tensor-core tile product D(8x8) += A(8x4) B(4x8), PTX fragment layout of
  // mma.sync.m8n8k4.f64: lane = 4 g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
  static OSC_HD void mma884(Var<double>& d0, Var<double>& d1, const Var<double>& a,
                            const Var<double>& b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0.v), "+d"(d1.v)
                 : "d"(a.v), "d"(b.v));
  }
};

#else  // host emulation of one warp

template <class T>
struct Var {
  T v[32];
  T& operator[](int l) { return v[l]; }
  const T& operator[](int l) const { return v[l]; }
};
// OSC_WARP_REVERSE runs the lanes of every body in the opposite order: a body in which one
// lane reads what another lane writes (a race on the GPU) then gives different results, which
// tests/test_host_core.py checks for
#ifdef OSC_WARP_REVERSE
#define OSC_LANES(l) for (int l = 31; l >= 0; --l)
#else
#define OSC_LANES(l) for (int l = 0; l < 32; ++l)
#endif

struct Warp {
  static void sync() {}
  static void xchg16(Var<double>& dst, const Var<double>& src) {
    double t[32];
    for (int l = 0; l < 32; ++l) t[l] = src.v[l ^ 16];
    for (int l = 0; l < 32; ++l) dst.v[l] = t[l];
  }
  static void group4(Var<double>& dst, const Var<double>& src, int r) {
    double t[32];
    for (int l = 0; l < 32; ++l) t[l] = src.v[(l & ~3) | r];
    for (int l = 0; l < 32; ++l) dst.v[l] = t[l];
  }
  static double bcast(const Var<double>& src, int from) { return src.v[from]; }
  static double sum(const Var<double>& a) {  // the device's butterfly order
    double v[32], t[32];
    for (int l = 0; l < 32; ++l) v[l] = a.v[l];
    for (int o = 16; o > 0; o >>= 1) {
      for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ o];
      for (int l = 0; l < 32; ++l) v[l] = t[l];
    }
    return v[0];
  }
  static double max(const Var<double>& a) {  // non-negative values; a NaN never wins
    double m = 0.0;
    for (int l = 0; l < 32; ++l) m = a.v[l] > m ? a.v[l] : m;
    return m;
  }
  static unsigned ballot(const Var<bool>& p) {
    unsigned b = 0;
    for (int l = 0; l < 32; ++l)
      if (p.v[l]) b |= 1u << l;
    return b;
  }
  template <int NQ>
  static void maxn(Var<double> (&m)[NQ], double* out, double* scratch, int) {
    for (int q = 0; q < NQ; ++q) out[q] = scratch[q] = max(m[q]);
  }
  static void max16(Var<double> (&m)[16], double* out, double* scratch, int) {
    maxn<16>(m, out, scratch, 0);
  }
  static void mma884(Var<double>& d0, Var<double>& d1, const Var<double>& a,
                     const Var<double>& b) {
    for (int l = 0; l < 32; ++l) {
      const int g = l >> 2, t = l & 3;
      double s0 = d0.v[l], s1 = d1.v[l];
      for (int k = 0; k < 4; ++k) {
        s0 = fma(a.v[4 * g + k], b.v[4 * (2 * t) + k], s0);
        s1 = fma(a.v[4 * g + k], b.v[4 * (2 * t + 1) + k], s1);
      }
      d0.v[l] = s0;
      d1.v[l] = s1;
    }
  }
};

#endif

}  // namespace osc
