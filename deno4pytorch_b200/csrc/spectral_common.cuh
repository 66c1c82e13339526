// spectral_common.cuh - shared definitions for the spectral-convolution kernels.
//
// Compiles two ways: with nvcc for sm_100a (the product) and, under -DDENO_EMULATE, with g++
// against tests/emu/cuda_emu.h so that the CPU test-suite can exercise the SIMT kernels'
// index arithmetic and launch geometry without a GPU (test infrastructure only).
#pragma once

#ifdef DENO_EMULATE
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <stdlib.h>
#include <utility>
#define DENO_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
namespace deno {
// Every kernel of the library can be launched with programmatic dependent launch (PDL, DENO_PDL=1): the next kernel's
// CTAs may be scheduled while this one is still running, run their input-independent prologue (TMEM allocation, barrier
// initialisation, constant / weight operand images) and then block in DENO_PDL_SYNC() until the predecessor has
// completed and flushed.  OFF by default: inside the CUDA-graph replay of the training step it measured 19 712 vs
// 19 735 samples/s without (round 1, profiles/r2w note in DESIGN.md) - the graph already removes the launch gaps and
// the kernels fill the SMs to their end - so the default launch stays the plain, fully serialised one.  Rules every kernel follows: (1) nothing the IMMEDIATE predecessor may write is read, and no
// global memory is written, before DENO_PDL_SYNC(); (2) DENO_PDL_SYNC() also releases the successor
// (launch_dependents), i.e. only after this kernel has seen its own predecessor complete - so completion is
// transitive and weights / constants written several kernels earlier are safe to read in the prologue.
// Without the attribute the wait returns immediately.
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DENO_PDL"); return e != nullptr && e[0] == '1'; }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace deno
#define DENO_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)deno::launch_pdl((kernel), (grid), (block), (smem), (stream), __VA_ARGS__)
#define DENO_PDL_SYNC() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
#endif

#include <stdint.h>

// Compiler-level scheduling fence: memory operations may not be moved across it.  Used after a batch of
// independent global loads so that all of them are in flight before the first one is consumed.
#ifdef DENO_EMULATE
#define DENO_SCHED_FENCE() do { } while (0)
#else
#define DENO_SCHED_FENCE() asm volatile("" ::: "memory")
#endif

namespace deno {

// Ordered (asm volatile) global loads: ptxas keeps them in program order, so a batch written before the
// arithmetic really is issued as a batch - the way to get memory-level parallelism in the small
// latency-bound kernels (mode_mix*, reductions), where plain loads get re-interleaved with their uses.
__device__ __forceinline__ float2 ldg_ordered(const float2* p) {
#ifdef DENO_EMULATE
  return *p;
#else
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ float ldg_ordered(const float* p) {
#ifdef DENO_EMULATE
  return *p;
#else
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#endif
}

// Activation ids mirror enum deno_activation in include/deno_spectral.h
// (table of /root/reference/Models/configs.py:17-19).
enum : int { ACT_IDENTITY = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_RELU = 3, ACT_TANH = 4, ACT_LEAKYRELU = 5 };

// exp / reciprocal used by the activations: the hardware approximations (2 ulp) on the GPU.
__device__ __forceinline__ float fast_exp(float x) {
#ifdef DENO_EMULATE
  return expf(x);
#else
  return __expf(x);
#endif
}
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef DENO_EMULATE
  return 1.0f / x;
#else
  return __fdividef(1.0f, x);
#endif
}

// erf-form GELU (nn.GELU() default) and its derivative in ~20 instructions: Phi(z) through the
// Abramowitz-Stegun 7.1.26 rational approximation of erf, whose exponential exp(-(z/sqrt2)^2) is the very
// exp(-z^2/2) the derivative's pdf needs.  Measured against fp64 on 2.2e6 points in [-8, 8]: relative L2
// error 6.1e-8 (y) / 9.1e-8 (y'), i.e. the same as PyTorch's own fp32 GELU (6.9e-8); erff() costs ~5x more.
__device__ __forceinline__ void gelu_fast(float z, float& y, float& d) {
  const float ax = fabsf(z) * 0.70710678118654752440f;
  const float e = fast_exp(-ax * ax);
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, z), 0.5f);
  y = z * cdf;
  d = fmaf(z * e, 0.39894228040143267794f, cdf);
}

// Two GELUs at once on the packed fp32x2 pipe of sm_100 (fma.rn.f32x2 and friends: two IEEE fp32 operations per issue
// slot).  Same formula as gelu_fast; only exp2 / rcp / abs / copysign stay scalar.  The epilogues that evaluate 32-64
// GELUs per thread and tile are issue-bound, so this is where their time goes.
__device__ __forceinline__ void gelu_fast2(float2 z, float2& y, float2& d) {
#ifdef DENO_EMULATE
  gelu_fast(z.x, y.x, d.x);
  gelu_fast(z.y, y.y, d.y);
#else
  const float2 ax = __fmul2_rn(make_float2(fabsf(z.x), fabsf(z.y)), make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 earg = __fmul2_rn(__fmul2_rn(ax, ax), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e, t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(earg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(earg.y));
  const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.0f, 1.0f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  // q = -poly(t): coefficients negated so that erf_abs = q * t * e + 1
  float2 q = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
  q = __ffma2_rn(q, t, make_float2(-1.421413741f, -1.421413741f));
  q = __ffma2_rn(q, t, make_float2(0.284496736f, 0.284496736f));
  q = __ffma2_rn(q, t, make_float2(-0.254829592f, -0.254829592f));
  const float2 erf_abs = __ffma2_rn(__fmul2_rn(q, t), e, make_float2(1.0f, 1.0f));
  const float2 erf_s = make_float2(copysignf(erf_abs.x, z.x), copysignf(erf_abs.y, z.y));
  const float2 cdf = __ffma2_rn(make_float2(0.5f, 0.5f), erf_s, make_float2(0.5f, 0.5f));
  y = __fmul2_rn(z, cdf);
  d = __ffma2_rn(__fmul2_rn(z, e), make_float2(0.39894228040143267794f, 0.39894228040143267794f), cdf);
#endif
}

__device__ __forceinline__ float act_apply(int act, float z) {
  switch (act) {
    case ACT_GELU: {
      float y, d;
      gelu_fast(z, y, d);
      return y;
    }
    case ACT_SILU: return z * fast_rcp(1.0f + fast_exp(-z));
    case ACT_RELU: return z > 0.0f ? z : 0.0f;
    case ACT_TANH: return tanhf(z);
    case ACT_LEAKYRELU: return z > 0.0f ? z : 0.01f * z;
    default: return z;
  }
}

__device__ __forceinline__ float act_grad(int act, float z) {
  switch (act) {
    case ACT_GELU: {
      float y, d;
      gelu_fast(z, y, d);
      return d;
    }
    case ACT_SILU: {
      const float s = fast_rcp(1.0f + fast_exp(-z));
      return s * (1.0f + z * (1.0f - s));
    }
    case ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
    case ACT_TANH: {
      const float t = tanhf(z);
      return 1.0f - t * t;
    }
    case ACT_LEAKYRELU: return z > 0.0f ? 1.0f : 0.01f;
    default: return 1.0f;
  }
}

// y = act(z) and d = act'(z) in one go (the forward epilogue saves d for the backward pass).
__device__ __forceinline__ void act_apply_grad(int act, float z, float& y, float& d) {
  switch (act) {
    case ACT_GELU:
      gelu_fast(z, y, d);
      return;
    case ACT_SILU: {
      const float s = fast_rcp(1.0f + fast_exp(-z));
      y = z * s;
      d = s * (1.0f + z * (1.0f - s));
      return;
    }
    case ACT_TANH: {
      const float t = tanhf(z);
      y = t;
      d = 1.0f - t * t;
      return;
    }
    default:
      y = act_apply(act, z);
      d = act_grad(act, z);
      return;
  }
}

// A "plane problem": the real tensor is seen as NB' = NBO*NBI batches of C channels of contiguous
// H x W planes (W fastest).  1-D layers use H = 1; 3-D layers (X,Y,Z) use planes (Y,Z) with
// NBI = X, so the same plane kernels serve every dimensionality.
struct PlaneGeom {
  int H, W;             // plane rows / cols
  int KH;               // packed modes along H (2*m, or 1 when H == 1)
  int MW;               // retained modes along W (half spectrum)
  int C;                // channels of the tensor this geometry addresses
  int NBO, NBI;         // batch = bo * NBI + bi
  long long sbo, sbi, sc;  // element strides of bo, bi, channel
};

__host__ __device__ __forceinline__ long long plane_base(const PlaneGeom& g, int bprime, int c) {
  int bo = bprime / g.NBI;
  int bi = bprime - bo * g.NBI;
  return (long long)bo * g.sbo + (long long)bi * g.sbi + (long long)c * g.sc;
}

__host__ __device__ __forceinline__ int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Maps a packed mode index to (weight block, offset inside the block); see oracle/dft_truth.py
// for the packed layout and spectral_layers.py:196-199,318-325 for the block order.
struct ModeDecode {
  int ndim;
  int m1, m2, m3;  // modes per axis (unused ones = 1)
  int M;           // total packed modes
  int Mblk;        // modes per weight block
};

__host__ __device__ __forceinline__ void decode_mode(const ModeDecode& md, int p, int& blk, int& off) {
  if (md.ndim == 1) {
    blk = 0;
    off = p;
  } else if (md.ndim == 2) {
    int k = p / md.m2, l = p - k * md.m2;
    blk = k >= md.m1 ? 1 : 0;
    off = (k - blk * md.m1) * md.m2 + l;
  } else {
    int l = p % md.m3;
    int t = p / md.m3;
    int k2 = t % (2 * md.m2);
    int k1 = t / (2 * md.m2);
    int b1 = k1 >= md.m1 ? 1 : 0, b2 = k2 >= md.m2 ? 1 : 0;
    blk = b1 + 2 * b2;
    off = ((k1 - b1 * md.m1) * md.m2 + (k2 - b2 * md.m2)) * md.m3 + l;
  }
}

// Byte offset of element (r, k) inside a tcgen05 K-major, no-swizzle operand image with `rows` rows:
// 128-byte core matrices (8 rows x 4 fp32), row groups contiguous (SBO = 128), LBO = 16 * rows.
__host__ __device__ __forceinline__ int kmajor_off(int r, int k, int rows) {
  return (r & 7) * 16 + (r >> 3) * 128 + (k & 3) * 4 + (k >> 2) * (16 * rows);
}

}  // namespace deno
