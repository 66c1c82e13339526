// afno_pointwise.cuh - the per-bin block-diagonal complex MLP of the adaptive Fourier layers
// (/root/reference/Models/fno/spectral_layers.py:377-405, :451-483, :527-561), fused:
//
//   forward   out = softshrink( W2^T act( W1^T xhat + b1 ) + b2 )          one read and one write of the spectrum
//   backward  gxhat = conj(W1) ( act'(.) (.) conj(W2) ( shrink'(.) (.) ghat ) )   + per-block gW1, gb1, gW2, gb2
//
// replacing, per direction, two einsum launches (each lowered to permute + clone + bmm), two broadcast bias adds,
// view_as_real / view_as_complex round trips, the activation, two zero-filled spectra with slice assignment and the
// soft-shrink: ~12 spectrum-sized passes in the reference, 2 here.  The hidden layer is never written to memory; the
// backward recomputes it from the saved input spectrum.
//
// Layout: spectra are [B][C][M] complex64, M bins fastest (what the analysis kernels write).  A CTA owns one channel
// block n and walks tiles of TP consecutive bins of one batch entry: thread = bin (coalesced along M), the block's
// bs x bs weights are staged once in shared memory and broadcast, four output channels per pass over the inputs.
// Weight gradients: the tile's operands stay in shared memory, thread = (in, out) pair sums over the tile's bins into
// registers across ALL tiles of the CTA, one partial row per CTA, reduced in a fixed order by bypass_wgrad_reduce
// (deterministic, no atomics).
//
// SIMT only (also compiled by the CPU emulator of the test-suite).  Not yet measured on hardware.
#pragma once
#include "spectral_common.cuh"

namespace deno {

constexpr int kAfnoFwdThreads = 128;
constexpr int kAfnoBwdThreads = 256;
constexpr int kAfnoBwdTasks = 17;      // (bs * bs + bs) / 256 rounded up for bs <= 64
constexpr int kAfnoMaxBlock = 64;

struct AfnoParams {
  const float2* xin;    // [B][C][M] input spectrum
  const float2* gin;    // backward: cotangent of the output spectrum [B][C][M]
  float2* out;          // forward: output spectrum; backward: cotangent of the input spectrum (may be null)
  const float2* w1;     // [NB][bs][bs]  (in, out)
  const float2* b1;     // [NB][bs]
  const float2* w2;
  const float2* b2;
  float* partial;       // backward: [gridDim.x][pitch] floats, row = [gW1 | gb1 | gW2 | gb2] of all blocks; may be null
  int B, C, NB, BS, M;
  int act;
  float lambda;
  int TP;               // bins per tile (<= blockDim.x)
  int tpb;              // tiles per batch entry
  int pitch;            // floats per partial row
};

// floats of one partial row and the offsets of its four sections
struct AfnoPartial { int w1, b1, w2, b2, total; };
__host__ __device__ inline AfnoPartial afno_partial(int NB, int BS) {
  AfnoPartial a;
  a.w1 = 0;
  a.b1 = a.w1 + NB * BS * BS * 2;
  a.w2 = a.b1 + NB * BS * 2;
  a.b2 = a.w2 + NB * BS * BS * 2;
  a.total = round_up(a.b2 + NB * BS * 2, 4);
  return a;
}

// shared memory (float2 units): W1, W2 [bs][BSp], b1, b2 [BSp], then `ntiles` operand tiles [bs][TP + 1]
__host__ __device__ inline size_t afno_smem_bytes(int BS, int TP, int ntiles) {
  const int BSp = round_up(BS, 4);
  return ((size_t)2 * BS * BSp + 2 * BSp + (size_t)ntiles * BS * (TP + 1)) * sizeof(float2);
}

__device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 acc) {       // acc + a * b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ float2 cfma_conj(float2 a, float2 b, float2 acc) {  // acc + conj(a) * b
  acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ float softshrink(float v, float lambda) {
  return v > lambda ? v - lambda : (v < -lambda ? v + lambda : 0.0f);
}

// Stages the block's weights: Ws [i][BSp] (columns beyond bs zero), bias [BSp].
__device__ __forceinline__ void afno_stage_weights(const AfnoParams& p, int n, float2* W1s, float2* W2s, float2* b1s,
                                                   float2* b2s, int BSp) {
  const int bs = p.BS, tid = threadIdx.x, nt = blockDim.x;
  const float2 zero = make_float2(0.f, 0.f);
  for (int e = tid; e < bs * BSp; e += nt) {
    const int i = e / BSp, o = e - i * BSp;
    W1s[e] = (o < bs) ? p.w1[((long long)n * bs + i) * bs + o] : zero;
    W2s[e] = (o < bs) ? p.w2[((long long)n * bs + i) * bs + o] : zero;
  }
  for (int o = tid; o < BSp; o += nt) {
    b1s[o] = (o < bs) ? p.b1[(long long)n * bs + o] : zero;
    b2s[o] = (o < bs) ? p.b2[(long long)n * bs + o] : zero;
  }
}

// dst[o][col] = f( bias[o] + sum_i src[i][col] * W[i][o] ) for all o, four outputs per pass; `f` is applied by the caller
// through the functor so the same loop serves both layers.
template <typename Epilogue>
__device__ __forceinline__ void afno_matvec(const float2* src, const float2* Ws, const float2* bias, int bs, int BSp,
                                            int pitch, int col, Epilogue epi) {
  for (int o0 = 0; o0 < bs; o0 += 4) {
    float2 a0 = bias[o0], a1 = bias[o0 + 1], a2 = bias[o0 + 2], a3 = bias[o0 + 3];
    for (int i = 0; i < bs; ++i) {
      const float2 v = src[i * pitch + col];
      const float2* w = Ws + i * BSp + o0;
      a0 = cfma(v, w[0], a0);
      a1 = cfma(v, w[1], a1);
      a2 = cfma(v, w[2], a2);
      a3 = cfma(v, w[3], a3);
    }
    epi(o0, a0);
    if (o0 + 1 < bs) epi(o0 + 1, a1);
    if (o0 + 2 < bs) epi(o0 + 2, a2);
    if (o0 + 3 < bs) epi(o0 + 3, a3);
  }
}

// dst[i] = sum_o conj(W[i][o]) * src[o][col]   (the transposed, conjugated product of the backward pass)
template <typename Epilogue>
__device__ __forceinline__ void afno_matvec_conj_t(const float2* src, const float2* Ws, int bs, int BSp, int pitch,
                                                   int col, Epilogue epi) {
  for (int i = 0; i < bs; ++i) {
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
    const float2* w = Ws + i * BSp;
    int o = 0;
    for (; o + 1 < bs; o += 2) {
      a = cfma_conj(w[o], src[o * pitch + col], a);
      b = cfma_conj(w[o + 1], src[(o + 1) * pitch + col], b);
    }
    if (o < bs) a = cfma_conj(w[o], src[o * pitch + col], a);
    epi(i, make_float2(a.x + b.x, a.y + b.y));
  }
}

// grid (chunks, NB), block kAfnoFwdThreads
__global__ void __launch_bounds__(kAfnoFwdThreads) afno_mlp_fwd_kernel(AfnoParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  const int bs = p.BS, BSp = round_up(bs, 4), TP = p.TP, pitch = TP + 1;
  float2* W1s = reinterpret_cast<float2*>(smem_raw);
  float2* W2s = W1s + bs * BSp;
  float2* b1s = W2s + bs * BSp;
  float2* b2s = b1s + BSp;
  float2* xs = b2s + BSp;               // [bs][pitch]
  float2* hs = xs + bs * pitch;         // [bs][pitch]
  const int n = blockIdx.y, tid = threadIdx.x;
  afno_stage_weights(p, n, W1s, W2s, b1s, b2s, BSp);
  __syncthreads();
  const int ntiles = p.B * p.tpb;
  const int act = p.act;
  const float lambda = p.lambda;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = t / p.tpb;
    const int q0 = (t - b * p.tpb) * TP;
    const int npos = min(TP, p.M - q0);
    if (tid < npos) {
      const long long base = ((long long)b * p.C + (long long)n * bs) * p.M + q0 + tid;
      for (int i = 0; i < bs; ++i) xs[i * pitch + tid] = p.xin[base + (long long)i * p.M];
      afno_matvec(xs, W1s, b1s, bs, BSp, pitch, tid, [&](int o, float2 v) {
        hs[o * pitch + tid] = make_float2(act_apply(act, v.x), act_apply(act, v.y));
      });
      afno_matvec(hs, W2s, b2s, bs, BSp, pitch, tid, [&](int o, float2 v) {
        p.out[base + (long long)o * p.M] = make_float2(softshrink(v.x, lambda), softshrink(v.y, lambda));
      });
    }
    // every thread touches only its own column of xs / hs: no barrier needed between tiles
  }
}

// grid (chunks, NB), block kAfnoBwdThreads.  Tiles: xs (input), hs (hidden), gs (masked output cotangent),
// ds (act' then the pre-activation cotangent).
__global__ void __launch_bounds__(kAfnoBwdThreads) afno_mlp_bwd_kernel(AfnoParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  const int bs = p.BS, BSp = round_up(bs, 4), TP = p.TP, pitch = TP + 1;
  float2* W1s = reinterpret_cast<float2*>(smem_raw);
  float2* W2s = W1s + bs * BSp;
  float2* b1s = W2s + bs * BSp;
  float2* b2s = b1s + BSp;
  float2* xs = b2s + BSp;
  float2* hs = xs + bs * pitch;
  float2* gs = hs + bs * pitch;
  float2* ds = gs + bs * pitch;
  const int n = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  afno_stage_weights(p, n, W1s, W2s, b1s, b2s, BSp);
  __syncthreads();
  const int ntiles = p.B * p.tpb;
  const int act = p.act;
  const float lambda = p.lambda;
  const bool want_w = p.partial != nullptr;
  const int ntask = bs * bs + bs;          // (in, out) pairs, then the bias entries
  float2 acc1[kAfnoBwdTasks], acc2[kAfnoBwdTasks];
#pragma unroll
  for (int k = 0; k < kAfnoBwdTasks; ++k) acc1[k] = acc2[k] = make_float2(0.f, 0.f);

  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = t / p.tpb;
    const int q0 = (t - b * p.tpb) * TP;
    const int npos = min(TP, p.M - q0);
    if (tid < npos) {
      const long long base = ((long long)b * p.C + (long long)n * bs) * p.M + q0 + tid;
      for (int i = 0; i < bs; ++i) {
        xs[i * pitch + tid] = p.xin[base + (long long)i * p.M];
        gs[i * pitch + tid] = p.gin[base + (long long)i * p.M];
      }
      // hidden layer again: h = act(pre), d = act'(pre)
      afno_matvec(xs, W1s, b1s, bs, BSp, pitch, tid, [&](int o, float2 v) {
        float hx, hy, dx, dy;
        act_apply_grad(act, v.x, hx, dx);
        act_apply_grad(act, v.y, hy, dy);
        hs[o * pitch + tid] = make_float2(hx, hy);
        ds[o * pitch + tid] = make_float2(dx, dy);
      });
      // output cotangent through the soft-shrink: passes where |pre-shrink value| > lambda
      afno_matvec(hs, W2s, b2s, bs, BSp, pitch, tid, [&](int o, float2 v) {
        float2 g = gs[o * pitch + tid];
        if (!(fabsf(v.x) > lambda)) g.x = 0.f;
        if (!(fabsf(v.y) > lambda)) g.y = 0.f;
        gs[o * pitch + tid] = g;
      });
      // hidden cotangent, through the activation (real and imaginary parts separately)
      afno_matvec_conj_t(gs, W2s, bs, BSp, pitch, tid, [&](int i, float2 v) {
        const float2 d = ds[i * pitch + tid];
        ds[i * pitch + tid] = make_float2(v.x * d.x, v.y * d.y);
      });
      if (p.out != nullptr) {
        afno_matvec_conj_t(ds, W1s, bs, BSp, pitch, tid, [&](int i, float2 v) { p.out[base + (long long)i * p.M] = v; });
      }
    } else if (tid < TP && want_w) {
      const float2 zero = make_float2(0.f, 0.f);
      for (int i = 0; i < bs; ++i) {
        xs[i * pitch + tid] = zero; hs[i * pitch + tid] = zero; gs[i * pitch + tid] = zero; ds[i * pitch + tid] = zero;
      }
    }
    if (want_w) {
      __syncthreads();
      // gW2[i][o] += sum_bins conj(h_i) go_o ;  gW1[i][o] += sum_bins conj(x_i) gpre_o ;  gb2[o] += go_o ; gb1[o] += gpre_o
#pragma unroll
      for (int k = 0; k < kAfnoBwdTasks; ++k) {
        const int task = tid + k * nt;
        if (task < bs * bs) {
          const int i = task / bs, o = task - i * bs;
          const float2* hi = hs + i * pitch;
          const float2* xi = xs + i * pitch;
          const float2* go = gs + o * pitch;
          const float2* dp = ds + o * pitch;
          float2 a2 = acc2[k], a1 = acc1[k];
          for (int c = 0; c < TP; ++c) {
            a2 = cfma_conj(hi[c], go[c], a2);
            a1 = cfma_conj(xi[c], dp[c], a1);
          }
          acc2[k] = a2; acc1[k] = a1;
        } else if (task < ntask) {
          const int o = task - bs * bs;
          const float2* go = gs + o * pitch;
          const float2* dp = ds + o * pitch;
          float2 a2 = acc2[k], a1 = acc1[k];
          for (int c = 0; c < TP; ++c) {
            a2.x += go[c].x; a2.y += go[c].y;
            a1.x += dp[c].x; a1.y += dp[c].y;
          }
          acc2[k] = a2; acc1[k] = a1;
        }
      }
      __syncthreads();
    }
  }
  if (want_w) {
    const AfnoPartial pl = afno_partial(p.NB, bs);
    float* row = p.partial + (long long)blockIdx.x * p.pitch;
#pragma unroll
    for (int k = 0; k < kAfnoBwdTasks; ++k) {
      const int task = tid + k * nt;
      if (task < bs * bs) {
        const int e = (n * bs * bs + task) * 2;
        row[pl.w1 + e] = acc1[k].x; row[pl.w1 + e + 1] = acc1[k].y;
        row[pl.w2 + e] = acc2[k].x; row[pl.w2 + e + 1] = acc2[k].y;
      } else if (task < ntask) {
        const int e = (n * bs + task - bs * bs) * 2;
        row[pl.b1 + e] = acc1[k].x; row[pl.b1 + e + 1] = acc1[k].y;
        row[pl.b2 + e] = acc2[k].x; row[pl.b2 + e + 1] = acc2[k].y;
      }
    }
    if (n == 0 && tid < pl.total - (pl.b2 + p.NB * bs * 2)) row[pl.b2 + p.NB * bs * 2 + tid] = 0.f;   // row padding
  }
}

}  // namespace deno
