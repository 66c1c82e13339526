// fno_pointwise.cuh - the point-wise ends of an FNO stack, fused (SURVEY.md section 8f "next" rows):
//
//   lift        x (B, *S, Kx) ++ grid (B, *S, nd)  --fc0-->  channel-first, right-padded h (B, C, *P)
//               replaces  cat + Linear + permute + F.pad          (/root/reference/Models/fno/FNOs.py:136-142)
//   projection  h (B, C, *P)  --crop, fc1, GELU, fc2-->  y (B, *S, O)
//               replaces  slice + permute + Linear + gelu + Linear (FNOs.py:147-152)
//
// and their backward passes.  The torch formulation moves every intermediate through HBM (the 128-wide
// hidden layer of the projection alone is 74 MB per pass for the Darcy config, saved twice for autograd);
// here a thread owns one grid point, keeps its C channel values in registers, walks the hidden units with the
// weights broadcast from shared memory, and nothing but h, y and the gradients touches memory.  The backward
// kernel recomputes the hidden layer instead of saving it.  Reductions over points (bias and fc2 weight
// gradients) go through shared memory in a fixed order and per-block partials (deterministic); the fc1 weight
// gradient, a long-K GEMM, reuses the tensor-core bypass_wgrad path on the materialised pre-activation gradient.
//
// SIMT only (also compiled by the CPU emulator of the test-suite).
#pragma once
#include "spectral_common.cuh"

namespace deno {

// Cropped (S) and right-padded (P) extents of up to three spatial axes, outermost first; unused axes are 1.
struct PointGeom {
  int sx, sy, sz;
  int px, py, pz;
  int npts;    // sx * sy * sz
  int nppts;   // px * py * pz
};

// Padded point index -> cropped point index; false inside the padding.
__device__ __forceinline__ bool padded_to_cropped(const PointGeom& g, int pp, int& pc) {
  const int r = pp / g.pz;
  const int z = pp - r * g.pz;
  int x = 0, y = r;
  if (g.px > 1) {                      // one integer division for 1-D / 2-D grids, two for 3-D
    x = r / g.py;
    y = r - x * g.py;
  }
  pc = (x * g.sy + y) * g.sz + z;
  return z < g.sz && y < g.sy && x < g.sx;
}

constexpr int kLiftMaxK = 16;       // steps * in_dim + ndim input features at most
constexpr int kLiftThreads = 256;

struct LiftParams {
  PointGeom g;
  int C, Kx, ND;         // width; features of x; coordinates of grid
  const float* x;        // (B, npts, Kx)
  const float* grid;     // (B, npts, ND)
  const float* w0;       // (C, Kx + ND)
  const float* b0;       // (C)
  float* h;              // (B, C, nppts)
  int chunks;            // ceil(nppts / kLiftThreads); gridDim.x = B * chunks
};

// KT = compile-time bound on the feature count K (2 .. 16): the per-channel dot product is KT FMAs with the features in
// registers.  With the bound fixed at 16 for every model the Darcy lift (K = 3) executed 72 warp instructions per
// 128-byte store (ncu: 12.9 M instructions, 19 us for a 22.6 MB write).
template <int KT>
__global__ void __launch_bounds__(kLiftThreads) fno_lift_kernel(LiftParams p) {
  DENO_PDL_SYNC();   // first kernel after the optimiser step: even the weights need the predecessor
  DENO_DYN_SMEM(smem_raw);
  float* w0s = reinterpret_cast<float*>(smem_raw);     // [C][K] then b0 [C]
  const int K = p.Kx + p.ND;
  float* b0s = w0s + p.C * K;
  for (int i = threadIdx.x; i < p.C * K; i += kLiftThreads) w0s[i] = p.w0[i];
  for (int i = threadIdx.x; i < p.C; i += kLiftThreads) b0s[i] = p.b0[i];
  __syncthreads();
  const int b = blockIdx.x / p.chunks;
  const int pp = (blockIdx.x % p.chunks) * kLiftThreads + threadIdx.x;
  if (pp >= p.g.nppts) return;
  int pc;
  const bool valid = padded_to_cropped(p.g, pp, pc);
  float v[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    float t = 0.f;
    if (valid && k < p.Kx) t = p.x[((long long)b * p.g.npts + pc) * p.Kx + k];
    else if (valid && k < K) t = p.grid[((long long)b * p.g.npts + pc) * p.ND + (k - p.Kx)];
    v[k] = t;
  }
  float* out = p.h + (long long)b * p.C * p.g.nppts + pp;
  if (!valid) {                                        // padding: zeros
    for (int c = 0; c < p.C; ++c) out[(long long)c * p.g.nppts] = 0.f;
    return;
  }
  const float* wr = w0s;
  for (int c = 0; c < p.C; ++c, wr += K) {
    float acc = b0s[c];
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) acc = fmaf(wr[k], v[k], acc);
    out[(long long)c * p.g.nppts] = acc;
  }
}

// Lift backward: gw0[c, k] = sum_pts g[b, c, pt] in[b, pt, k], gb0[c] = sum_pts g[b, c, pt] (as one extra
// "feature" that is 1 on valid points), optionally gx[b, pt, k] = sum_c g[b, c, pt] w0[c, k] for the x features
// (autoregressive roll-outs feed predictions back in).  One block per 256 padded points: the gradient tile goes
// through shared memory [C][257] and thread (c, k) sums its pair over the tile; per-block partials are summed by
// bypass_wgrad_reduce_kernel.
struct LiftBwdParams {
  PointGeom g;
  int C, Kx, ND;
  const float* gh;       // (B, C, nppts)
  const float* x;
  const float* grid;
  const float* w0;
  float* gx;             // (B, npts, Kx) or null
  float* partial;        // [gridDim.x][C * K (gw0) | C (gb0)]
  int chunks;
};

__host__ __device__ inline size_t lift_bwd_smem_bytes(int C, int K) {
  return ((size_t)C * (kLiftThreads + 1) + (size_t)kLiftThreads * (K + 1) + (size_t)C * K) * sizeof(float);
}

__global__ void __launch_bounds__(kLiftThreads) fno_lift_bwd_kernel(LiftBwdParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  const int K = p.Kx + p.ND, K1 = K + 1;
  float* s_g = reinterpret_cast<float*>(smem_raw);                 // [C][257]
  float* s_in = s_g + p.C * (kLiftThreads + 1);                    // [256][K + 1]
  float* w0s = s_in + kLiftThreads * K1;                           // [C][K]
  const int tid = threadIdx.x;
  for (int i = tid; i < p.C * K; i += kLiftThreads) w0s[i] = p.w0[i];
  const int b = blockIdx.x / p.chunks;
  const int pp = (blockIdx.x % p.chunks) * kLiftThreads + tid;
  int pc = 0;
  const bool valid = pp < p.g.nppts && padded_to_cropped(p.g, pp, pc);
  for (int k = 0; k < K1; ++k) {
    float t = 0.f;
    if (valid) {
      if (k < p.Kx) t = p.x[((long long)b * p.g.npts + pc) * p.Kx + k];
      else if (k < K) t = p.grid[((long long)b * p.g.npts + pc) * p.ND + (k - p.Kx)];
      else t = 1.0f;
    }
    s_in[tid * K1 + k] = t;
  }
  __syncthreads();                                                  // w0s visible
  float gxa[kLiftMaxK];
#pragma unroll
  for (int k = 0; k < kLiftMaxK; ++k) gxa[k] = 0.f;
  const float* src = p.gh + (long long)b * p.C * p.g.nppts + pp;
  for (int c = 0; c < p.C; ++c) {
    const float v = valid ? src[(long long)c * p.g.nppts] : 0.f;
    s_g[c * (kLiftThreads + 1) + tid] = v;
    if (p.gx != nullptr) {
#pragma unroll
      for (int k = 0; k < kLiftMaxK; ++k)
        if (k < p.Kx) gxa[k] = fmaf(v, w0s[c * K + k], gxa[k]);
    }
  }
  if (p.gx != nullptr && valid) {
#pragma unroll
    for (int k = 0; k < kLiftMaxK; ++k)
      if (k < p.Kx) p.gx[((long long)b * p.g.npts + pc) * p.Kx + k] = gxa[k];
  }
  __syncthreads();
  float* out = p.partial + (long long)blockIdx.x * p.C * K1;
  for (int pair = tid; pair < p.C * K1; pair += kLiftThreads) {
    const int c = pair / K1, k = pair - c * K1;
    const float* gr = s_g + c * (kLiftThreads + 1);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int pt = 0; pt < kLiftThreads; pt += 4) {
      a0 = fmaf(gr[pt], s_in[pt * K1 + k], a0);
      a1 = fmaf(gr[pt + 1], s_in[(pt + 1) * K1 + k], a1);
      a2 = fmaf(gr[pt + 2], s_in[(pt + 2) * K1 + k], a2);
      a3 = fmaf(gr[pt + 3], s_in[(pt + 3) * K1 + k], a3);
    }
    out[k < K ? c * K + k : p.C * K + c] = (a0 + a1) + (a2 + a3);   // [gw0 (C, K) | gb0 (C)]
  }
}

// ------------------------------------------------------------------------------------------------------
// Projection.  One thread per padded grid point (threads in the padding only write zeros in the backward).
// ------------------------------------------------------------------------------------------------------
constexpr int kProjThreads = 128;
constexpr int kProjMaxO = 4;        // out_dim of the fused path
constexpr int kProjJC = 16;         // hidden units per shared-memory reduction round of the backward

struct ProjParams {
  PointGeom g;
  int C, Hd, O;
  const float* h;        // (B, C, nppts)
  const float* w1;       // (Hd, C)
  const float* b1;       // (Hd)
  const float* w2;       // (O, Hd)
  const float* b2;       // (O)
  float* y;              // forward: (B, npts, O)
  const float* gy;       // backward: (B, npts, O)
  float* gh;             // backward: (B, C, nppts), zero in the padding
  float* gt;             // backward: (B, Hd, nppts) gradient at the fc1 output, zero in the padding
  float* partial;        // backward: [gridDim.x][Hd + O * Hd + O]  (gb1 | gw2 | gb2)
  int chunks;            // ceil(nppts / (kProjThreads * PT)); gridDim.x = B * chunks
};

__host__ __device__ inline size_t proj_weights_floats(int C, int Hd, int O) {
  return (size_t)Hd * C + Hd + (size_t)O * Hd + O;
}
__host__ __device__ inline size_t proj_fwd_smem_bytes(int C, int Hd, int O) { return proj_weights_floats(C, Hd, O) * sizeof(float); }
__host__ __device__ inline size_t proj_bwd_smem_bytes(int C, int Hd, int O) {
  return (proj_weights_floats(C, Hd, O) + (size_t)(1 + O) * kProjJC * (kProjThreads + 1) + (size_t)(1 + O) * kProjJC * 4) * sizeof(float);
}

__device__ __forceinline__ void proj_load_weights(const ProjParams& p, float* w1s, float* b1s, float* w2s, float* b2s) {
  const float4* s4 = reinterpret_cast<const float4*>(p.w1);        // C % 4 == 0: Hd * C is a multiple of 4
  float4* d4 = reinterpret_cast<float4*>(w1s);
  for (int i = threadIdx.x; i < p.Hd * p.C / 4; i += kProjThreads) d4[i] = s4[i];
  for (int i = threadIdx.x; i < p.Hd; i += kProjThreads) b1s[i] = p.b1[i];
  for (int i = threadIdx.x; i < p.O * p.Hd; i += kProjThreads) w2s[i] = p.w2[i];
  for (int i = threadIdx.x; i < p.O; i += kProjThreads) b2s[i] = p.b2 != nullptr ? p.b2[i] : 0.f;   // unused by the backward
}

// A shared-memory broadcast delivers one weight per lane per LSU cycle and the LSU serves all four schedulers,
// so a weight has to feed several FMAs: every thread owns PT points (PT * kProjThreads apart, so loads and stores
// stay coalesced) and keeps the weight row of the current hidden unit in registers for all of them.
template <int C, int PT>
__global__ void __launch_bounds__(kProjThreads) fno_proj_fwd_kernel(ProjParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* w1s = reinterpret_cast<float*>(smem_raw);     // [Hd][C]
  float* b1s = w1s + p.Hd * C;
  float* w2s = b1s + p.Hd;                             // [O][Hd]
  float* b2s = w2s + p.O * p.Hd;
  proj_load_weights(p, w1s, b1s, w2s, b2s);
  const int b = blockIdx.x / p.chunks;
  const int pp0 = (blockIdx.x % p.chunks) * (kProjThreads * PT) + threadIdx.x;
  int pc[PT];
  bool valid[PT];
  float hr[PT][C];
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    const int pp = pp0 + i * kProjThreads;
    pc[i] = 0;
    valid[i] = pp < p.g.nppts && padded_to_cropped(p.g, pp, pc[i]);
    const float* src = p.h + (long long)b * C * p.g.nppts + pp;
#pragma unroll
    for (int c = 0; c < C; ++c) hr[i][c] = valid[i] ? __ldg(src + (long long)c * p.g.nppts) : 0.f;
  }
  __syncthreads();
  float yo[PT][kProjMaxO];
#pragma unroll
  for (int i = 0; i < PT; ++i)
#pragma unroll
    for (int o = 0; o < kProjMaxO; ++o) yo[i][o] = 0.f;
  for (int j = 0; j < p.Hd; ++j) {
    const float4* wr = reinterpret_cast<const float4*>(w1s + j * C);
    float4 wv[C / 4];
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) wv[c4] = wr[c4];
    const float bj = b1s[j];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      float t0 = bj, t1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        t0 = fmaf(wv[c4].x, hr[i][4 * c4], t0);
        t1 = fmaf(wv[c4].y, hr[i][4 * c4 + 1], t1);
        t0 = fmaf(wv[c4].z, hr[i][4 * c4 + 2], t0);
        t1 = fmaf(wv[c4].w, hr[i][4 * c4 + 3], t1);
      }
      float a, da;
      gelu_fast(t0 + t1, a, da);
#pragma unroll
      for (int o = 0; o < kProjMaxO; ++o)
        if (o < p.O) yo[i][o] = fmaf(w2s[o * p.Hd + j], a, yo[i][o]);
    }
  }
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    if (!valid[i]) continue;
    float* dst = p.y + ((long long)b * p.g.npts + pc[i]) * p.O;
#pragma unroll
    for (int o = 0; o < kProjMaxO; ++o)
      if (o < p.O) dst[o] = yo[i][o] + b2s[o];
  }
}

template <int C, int PT>
__global__ void __launch_bounds__(kProjThreads) fno_proj_bwd_kernel(ProjParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  const int R = (1 + p.O) * kProjJC;                   // reduction rows per round: gt, then gy_o * a per output
  float* w1s = reinterpret_cast<float*>(smem_raw);
  float* b1s = w1s + p.Hd * C;
  float* w2s = b1s + p.Hd;
  float* b2s = w2s + p.O * p.Hd;
  float* s_red = b2s + p.O;                            // [R][129]
  float* s_q = s_red + R * (kProjThreads + 1);         // [R][4]
  proj_load_weights(p, w1s, b1s, w2s, b2s);
  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.chunks;
  const int pp0 = (blockIdx.x % p.chunks) * (kProjThreads * PT) + tid;
  bool inside[PT];
  float hr[PT][C], gha[PT][C], gyr[PT][kProjMaxO];
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    const int pp = pp0 + i * kProjThreads;
    inside[i] = pp < p.g.nppts;
    int pc = 0;
    const bool valid = inside[i] && padded_to_cropped(p.g, pp, pc);
    const float* src = p.h + (long long)b * C * p.g.nppts + pp;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      hr[i][c] = valid ? __ldg(src + (long long)c * p.g.nppts) : 0.f;
      gha[i][c] = 0.f;
    }
#pragma unroll
    for (int o = 0; o < kProjMaxO; ++o)
      gyr[i][o] = (valid && o < p.O) ? __ldg(p.gy + ((long long)b * p.g.npts + pc) * p.O + o) : 0.f;
  }
  float* out = p.partial + (long long)blockIdx.x * (p.Hd + p.O * p.Hd + p.O);
  float* gt_dst = p.gt + (long long)b * p.Hd * p.g.nppts + pp0;
  __syncthreads();
  for (int j0 = 0; j0 < p.Hd; j0 += kProjJC) {
    for (int jj = 0; jj < kProjJC; ++jj) {
      const int j = j0 + jj;
      float gsum = 0.f, asum[kProjMaxO];
#pragma unroll
      for (int o = 0; o < kProjMaxO; ++o) asum[o] = 0.f;
      if (j < p.Hd) {
        const float4* wr = reinterpret_cast<const float4*>(w1s + j * C);
        float4 wv[C / 4];
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) wv[c4] = wr[c4];
        const float bj = b1s[j];
#pragma unroll
        for (int i = 0; i < PT; ++i) {
          float t0 = bj, t1 = 0.f;
#pragma unroll
          for (int c4 = 0; c4 < C / 4; ++c4) {
            t0 = fmaf(wv[c4].x, hr[i][4 * c4], t0);
            t1 = fmaf(wv[c4].y, hr[i][4 * c4 + 1], t1);
            t0 = fmaf(wv[c4].z, hr[i][4 * c4 + 2], t0);
            t1 = fmaf(wv[c4].w, hr[i][4 * c4 + 3], t1);
          }
          float av, da;
          gelu_fast(t0 + t1, av, da);
          float ga = 0.f;
#pragma unroll
          for (int o = 0; o < kProjMaxO; ++o)
            if (o < p.O) {
              ga = fmaf(gyr[i][o], w2s[o * p.Hd + j], ga);
              asum[o] = fmaf(gyr[i][o], av, asum[o]);
            }
          const float gtv = ga * da;                     // zero for padding threads: their gy is zero
          gsum += gtv;
#pragma unroll
          for (int c4 = 0; c4 < C / 4; ++c4) {
            gha[i][4 * c4] = fmaf(wv[c4].x, gtv, gha[i][4 * c4]);
            gha[i][4 * c4 + 1] = fmaf(wv[c4].y, gtv, gha[i][4 * c4 + 1]);
            gha[i][4 * c4 + 2] = fmaf(wv[c4].z, gtv, gha[i][4 * c4 + 2]);
            gha[i][4 * c4 + 3] = fmaf(wv[c4].w, gtv, gha[i][4 * c4 + 3]);
          }
          if (inside[i]) gt_dst[(long long)j * p.g.nppts + i * kProjThreads] = gtv;
        }
      }
      s_red[jj * (kProjThreads + 1) + tid] = gsum;
#pragma unroll
      for (int o = 0; o < kProjMaxO; ++o)
        if (o < p.O) s_red[((1 + o) * kProjJC + jj) * (kProjThreads + 1) + tid] = asum[o];
    }
    __syncthreads();
    for (int item = tid; item < R * 4; item += kProjThreads) {      // quarter sums, fixed order
      const int qt = item / R, row = item - qt * R;                 // lanes walk rows: conflict-free with the odd pitch
      const float* r = s_red + row * (kProjThreads + 1) + qt * (kProjThreads / 4);
      float s0 = 0.f, s1 = 0.f;
      for (int k = 0; k < kProjThreads / 4; k += 2) { s0 += r[k]; s1 += r[k + 1]; }
      s_q[row * 4 + qt] = s0 + s1;
    }
    __syncthreads();
    for (int row = tid; row < R; row += kProjThreads) {
      const int which = row / kProjJC, j = j0 + row % kProjJC;     // 0: gb1, 1 + o: gw2[o]
      if (j < p.Hd) out[which * p.Hd + j] = (s_q[row * 4] + s_q[row * 4 + 1]) + (s_q[row * 4 + 2] + s_q[row * 4 + 3]);
    }
  }
  // gb2[o] = sum_pts gy[pt, o]: one more round through the same buffers
  __syncthreads();
#pragma unroll
  for (int o = 0; o < kProjMaxO; ++o)
    if (o < p.O) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PT; ++i) s += gyr[i][o];
      s_red[o * (kProjThreads + 1) + tid] = s;
    }
  __syncthreads();
  if (tid < p.O) {
    const float* r = s_red + tid * (kProjThreads + 1);
    float s = 0.f;
    for (int k = 0; k < kProjThreads; ++k) s += r[k];
    out[p.Hd + p.O * p.Hd + tid] = s;
  }
#pragma unroll
  for (int i = 0; i < PT; ++i) {
    if (!inside[i]) continue;
    float* dst = p.gh + (long long)b * C * p.g.nppts + pp0 + i * kProjThreads;
#pragma unroll
    for (int c = 0; c < C; ++c) dst[(long long)c * p.g.nppts] = gha[i][c];
  }
}

// points per thread: as many as the register file allows next to the C channel values (and C gradients)
__host__ __device__ constexpr int proj_fwd_pt(int C) { return C <= 32 ? 4 : (C <= 48 ? 2 : 1); }
__host__ __device__ constexpr int proj_bwd_pt(int C) { return C <= 32 ? 2 : 1; }

}  // namespace deno
