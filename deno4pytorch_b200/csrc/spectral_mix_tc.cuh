// spectral_mix_tc.cuh - the per-mode complex channel mix on the tcgen05 tensor cores (sm_100a only), for layers of
// width >= 64 where it is a real dense GEMM (north_star item 2; reference: compl_mul1d/2d/3d,
// /root/reference/Models/fno/spectral_layers.py:73-82, 172-181, 287-297 and their autograd adjoints).
//
// Per retained mode m the forward product  Y^[b, o] = sum_i X^[b, i] W[i, o]  (complex) is ONE real GEMM with the real
// and imaginary parts STACKED along M and N, the batch on N:
//     A (M = 128, K = Cin)  rows  r <  T : Re W[., o]        rows T + r : Im W[., o]          (T = 64 produced channels)
//     B (N,       K = Cin)  rows  b      : Re X^[b, .]       rows Bp + b : Im X^[b, .]        (Bp = batch rounded up to 8)
//     D[r, n] = sum_k A[r, k] B[n, k]   ->   (Wr Xr)[o, b], (Wr Xi)[o, b], (Wi Xr)[o, b], (Wi Xi)[o, b] in one instruction
//     Re Y^ = Wr Xr - Wi Xi,  Im Y^ = Wr Xi + Wi Xr        (conjugated weights: signs of the Wi terms flipped)
// fp32-level accuracy by the 3xTF32 split with the hi and lo parts of B stacked along N as well (rows 2Bp.. = lo parts):
//     D  = A_hi x [B_hi | B_lo]   (N = 4 Bp)      D[:, 0 : 2Bp] += A_lo x B_hi   (N = 2 Bp)      T = D[:, n] + D[:, 2Bp + n]
// The rows T + r live in other TMEM lanes than the rows r, so the last combination goes through shared memory.
//
// The reference's weight layout is mode-INNERMOST ((Cin, Cout, m1, m2) complex64): the Cin x Cout matrix of one mode is
// a gather of 8-byte elements 8 * Mblk bytes apart.  A CTA therefore takes TM = 2 adjacent modes (one 16-byte load per
// (i, o) pair, both modes) when the shared-memory budget allows and TM = 1 otherwise (width 128); the other half of each
// 32-byte sector is read by the neighbouring CTA in the same wave and served by L2.
//
// mode_mix_tc_kernel      forward mix and the input-gradient mix  G^x[b, i] = sum_o conj(W[i, o]) G^[b, o]
// mode_mix_wgrad_tc_kernel  gW[i, o] = sum_b conj(X^[b, i]) G^[b, o]: M = stacked X^ (rows i), N = stacked G^ (rows o), K = batch
#pragma once
#include "spectral_tc.cuh"

namespace deno {
namespace tc {

constexpr int kMixTcThreads = 256;
constexpr int kMixTcTile = 64;          // channels per stacked operand half (2 * 64 = the M = 128 rows / N <= 256 columns)

struct MixTcParams {
  const float2* in;        // [B][CK][M]   contracted-channel spectrum
  float2* out;             // [B][CR][M]   produced-channel spectrum
  const float2* w[4];      // weight blocks
  ModeDecode md;
  int B, CK, CR;           // batch, contracted channels, produced channels
  long long w_sck, w_scr;  // complex-element strides of the contracted / produced channel inside a block
  int conj_w;
  int TM;                  // modes per CTA (1 or 2)
  int Bp, CKp;             // batch rounded up to 8, contracted channels rounded up to 8
  int ntile;               // produced-channel tiles of 64
};

struct MixTcSmem { int a, b, xch, total; int a_one, b_one, xch_pitch; };

__host__ __device__ inline MixTcSmem mix_tc_smem(int TM, int CKp, int Bp) {
  MixTcSmem m;
  m.a_one = 128 * CKp * 4;                 // one (hi or lo) A image of one mode
  m.b_one = 4 * Bp * CKp * 4;              // the stacked [Xr | Xi | Xr lo | Xi lo] B image of one mode
  m.xch_pitch = 2 * Bp + 1;
  int off = 0;
  m.a = off;   off += TM * 2 * m.a_one;    // [mode][hi | lo]
  m.b = off;   off += TM * m.b_one;
  m.xch = off; off += kMixTcTile * m.xch_pitch * 4;
  m.total = off + 2048;                    // M = 128 descriptors read 16 row groups of every core-matrix column
  return m;
}

__global__ void __launch_bounds__(kMixTcThreads, 1) mode_mix_tc_kernel(MixTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_mma;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TM = p.TM, Bp = p.Bp, CKp = p.CKp, B = p.B, CK = p.CK;
  const int NB = 4 * Bp, NH = 2 * Bp;
  const MixTcSmem so = mix_tc_smem(TM, CKp, Bp);
  const int M = p.md.M;
  const int tile = (int)blockIdx.x % p.ntile;
  const int pm0 = ((int)blockIdx.x / p.ntile) * TM;            // first packed mode of this CTA
  const int cr0 = tile * kMixTcTile;
  const int crn = min(kMixTcTile, p.CR - cr0);                 // produced channels of this tile
  uint32_t ncols = 32;
  while ((int)ncols < TM * NB) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  int blk, off0;
  decode_mode(p.md, pm0, blk, off0);
  const float2* wb = blk == 0 ? p.w[0] : (blk == 1 ? p.w[1] : (blk == 2 ? p.w[2] : p.w[3]));
  DENO_PDL_SYNC();

  // ---- A images: W of this tile's produced channels, both modes from one 16-byte load per (ck, cr) pair -------------
  // A warp covers an 8 (cr) x 4 (ck) patch per step: its 32 stores fill 128 contiguous bytes of the image (no conflicts).
  {
    const int npatch = (kMixTcTile / 8) * (CKp / 4);
    const int cr_l = lane & 7, ck_l = lane >> 3;
    for (int pt0 = warp; pt0 < npatch; pt0 += 8 * (kMixTcThreads / 32)) {
      float4 v[8];   // eight loads in flight per thread: the prologue is latency-bound
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        const int cr = (pt % (kMixTcTile / 8)) * 8 + cr_l, ck = (pt / (kMixTcTile / 8)) * 4 + ck_l;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt < npatch && cr < crn && ck < CK) {
          const float2* src = wb + (long long)ck * p.w_sck + (long long)(cr0 + cr) * p.w_scr + off0;
          if (TM == 2) {
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src));
          } else {
            const float2 t = ldg_ordered(src);
            v[u].x = t.x; v[u].y = t.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        if (pt < npatch) {
          const int cr = (pt % (kMixTcTile / 8)) * 8 + cr_l, ck = (pt / (kMixTcTile / 8)) * 4 + ck_l;
          const int o_re = kmajor_off(cr, ck, 128), o_im = kmajor_off(kMixTcTile + cr, ck, 128);
          const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          for (int t = 0; t < TM; ++t) {
            float h, l;
            unsigned char* ah = smem + so.a + t * 2 * so.a_one;
            split_tf32(vals[2 * t], h, l);
            *reinterpret_cast<float*>(ah + o_re) = h;
            *reinterpret_cast<float*>(ah + so.a_one + o_re) = l;
            split_tf32(vals[2 * t + 1], h, l);
            *reinterpret_cast<float*>(ah + o_im) = h;
            *reinterpret_cast<float*>(ah + so.a_one + o_im) = l;
          }
        }
      }
    }
  }
  // ---- B images: the contracted-channel spectrum of every batch entry (rows b >= B and channels >= CK are zeros) ----
  {
    const int npatch = (Bp / 8) * (CKp / 4);
    const int b_l = lane & 7, ck_l = lane >> 3;
    for (int pt0 = warp; pt0 < npatch; pt0 += 4 * (kMixTcThreads / 32)) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        const int b = (pt % (Bp / 8)) * 8 + b_l, ck = (pt / (Bp / 8)) * 4 + ck_l;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt < npatch && b < B && ck < CK) {
          const float2* src = p.in + ((long long)b * CK + ck) * M + pm0;
          if (TM == 2) {
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src));
          } else {
            const float2 t = ldg_ordered(src);
            v[u].x = t.x; v[u].y = t.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        if (pt < npatch) {
          const int b = (pt % (Bp / 8)) * 8 + b_l, ck = (pt / (Bp / 8)) * 4 + ck_l;
          const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          for (int t = 0; t < TM; ++t) {
            float h, l;
            unsigned char* bi = smem + so.b + t * so.b_one;
            split_tf32(vals[2 * t], h, l);
            *reinterpret_cast<float*>(bi + kmajor_off(b, ck, NB)) = h;
            *reinterpret_cast<float*>(bi + kmajor_off(NH + b, ck, NB)) = l;
            split_tf32(vals[2 * t + 1], h, l);
            *reinterpret_cast<float*>(bi + kmajor_off(Bp + b, ck, NB)) = h;
            *reinterpret_cast<float*>(bi + kmajor_off(NH + Bp + b, ck, NB)) = l;
          }
        }
      }
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  // ---- MMAs: one elected lane ----
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc_all = make_idesc_tf32(128, NB), idesc_hi = make_idesc_tf32(128, NH);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t lbo_a = 16 * 128, lbo_b = 16 * (uint32_t)NB;
    const uint64_t inc_a = (uint64_t)((2 * lbo_a) >> 4), inc_b = (uint64_t)((2 * lbo_b) >> 4);
    for (int t = 0; t < TM; ++t) {
      uint64_t ah = make_desc(sbase + so.a + t * 2 * so.a_one, lbo_a, 128);
      uint64_t al = make_desc(sbase + so.a + t * 2 * so.a_one + so.a_one, lbo_a, 128);
      uint64_t bb = make_desc(sbase + so.b + t * so.b_one, lbo_b, 128);
      const uint32_t d = tmem + (uint32_t)(t * NB);
      for (int s = 0; s < CKp / 8; ++s) {
        if (leader) {
          mma_tf32(d, ah, bb, idesc_all, s > 0 ? 1u : 0u);   // hi x [hi | lo]
          mma_tf32(d, al, bb, idesc_hi, 1u);                 // lo x hi
        }
        ah += inc_a; al += inc_a; bb += inc_b;
      }
    }
    if (leader) mma_commit(&bar_mma);
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();

  // ---- epilogue: warps 0-3 own the TMEM lanes; rows T + r hand their values to rows r through shared memory --------
  float* xch = reinterpret_cast<float*>(smem + so.xch);
  const int row = (warp & 3) * 32 + lane;
  const bool is_re = warp < 4 && row < crn, is_im = warp < 4 && row >= kMixTcTile && row < kMixTcTile + crn;
  const float sgn = p.conj_w ? -1.0f : 1.0f;
  for (int b0 = 0; b0 < Bp; b0 += 32) {                        // batch in passes of up to 32 entries
    float outv[2][64];                                         // [mode][re(b) | im(b)]
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t < TM) {
        float tr[32], ti[32];                                  // T[row, b0 + j] (Xr part), T[row, Bp + b0 + j] (Xi part)
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * NB);
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
          if (warp < 4 && b0 + j0 < Bp) {                      // warp-uniform: tcgen05.ld is warp-collective
            float a0[8], a1[8], c0[8], c1[8];
            tmem_ld8(ta + (uint32_t)(b0 + j0), a0);
            tmem_ld8(ta + (uint32_t)(NH + b0 + j0), a1);
            tmem_ld8(ta + (uint32_t)(Bp + b0 + j0), c0);
            tmem_ld8(ta + (uint32_t)(NH + Bp + b0 + j0), c1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { tr[j0 + j] = a0[j] + a1[j]; ti[j0 + j] = c0[j] + c1[j]; }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { tr[j0 + j] = 0.f; ti[j0 + j] = 0.f; }
          }
        }
        if (is_im) {
          float* dst = xch + (row - kMixTcTile) * so.xch_pitch + b0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (b0 + j < Bp) { dst[j] = tr[j]; dst[Bp + j] = ti[j]; }               // (Wi Xr)[b], (Wi Xi)[b]
        }
        __syncthreads();
        if (is_re) {
          const float* src = xch + row * so.xch_pitch + b0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (b0 + j < Bp) {
              outv[t][j] = tr[j] - sgn * src[Bp + j];            // Re: Wr Xr -/+ Wi Xi
              outv[t][32 + j] = ti[j] + sgn * src[j];            // Im: Wr Xi +/- Wi Xr
            }
          }
        }
        __syncthreads();
      }
    }
    if (is_re) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int b = b0 + j;
        if (b < B) {
          float2* dst = p.out + ((long long)b * p.CR + cr0 + row) * M + pm0;
          if (TM == 2) *reinterpret_cast<float4*>(dst) = make_float4(outv[0][j], outv[0][32 + j], outv[1][j], outv[1][32 + j]);
          else *dst = make_float2(outv[0][j], outv[0][32 + j]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// ----------------------------------------------------------------------------------------------
// gW[i, o, m] = sum_b conj(X^[b, i, m]) G^[b, o, m]
//   A rows i: Re X^[., i], rows T + i: Im X^[., i] (K = batch);  B rows o: Re G^, T + o: Im G^, then the lo parts
//   Re gW = Xr Gr + Xi Gi = T[i, o] + T[T + i, T + o]        Im gW = Xr Gi - Xi Gr = T[i, T + o] - T[T + i, o]
// One CTA = (TM modes, 64 input channels, 64 output channels).
// ----------------------------------------------------------------------------------------------
struct MixWgradTcParams {
  const float2* xhat;      // [B][CI][M]
  const float2* ghat;      // [B][CO][M]
  float* glb;              // nullable: bypass bias gradient from the DC mode, glb[o] = S * sum_b Re G^[b, o, 0]
  float s_total;
  float2* gw[4];
  ModeDecode md;
  int B, CI, CO;
  int TM, Kp;              // modes per CTA, batch rounded up to 8
  int nti, nto;            // channel tiles of 64
};

struct MixWgradTcSmem { int a, b, xch, total; int a_one, b_one, xch_pitch; };

__host__ __device__ inline MixWgradTcSmem mix_wgrad_tc_smem(int TM, int Kp) {
  MixWgradTcSmem m;
  m.a_one = 128 * Kp * 4;
  m.b_one = 4 * kMixTcTile * Kp * 4;
  m.xch_pitch = 2 * kMixTcTile + 1;
  int off = 0;
  m.a = off;   off += TM * 2 * m.a_one;
  m.b = off;   off += TM * m.b_one;
  m.xch = off; off += kMixTcTile * m.xch_pitch * 4;
  m.total = off + 2048;
  return m;
}

__global__ void __launch_bounds__(kMixTcThreads, 1) mode_mix_wgrad_tc_kernel(MixWgradTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_mma;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int T = kMixTcTile, NB = 4 * T, NH = 2 * T;
  const int TM = p.TM, Kp = p.Kp, B = p.B, M = p.md.M;
  const MixWgradTcSmem so = mix_wgrad_tc_smem(TM, Kp);
  int bid = blockIdx.x;
  const int to = bid % p.nto; bid /= p.nto;
  const int ti = bid % p.nti; bid /= p.nti;
  const int pm0 = bid * TM;
  const int i0 = ti * T, o0 = to * T;
  const int ni = min(T, p.CI - i0), no = min(T, p.CO - o0);
  uint32_t ncols = 32;
  while ((int)ncols < TM * NB) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
  }
  DENO_PDL_SYNC();

  // ---- operand images: element (channel row, k = batch entry); a warp covers an 8 (channel) x 4 (batch) patch ----
  auto stage = [&](const float2* src_base, int C, int c0, int nc, bool is_a) {
    const int npatch = (T / 8) * (Kp / 4);
    const int c_l = lane & 7, b_l = lane >> 3;
    for (int pt0 = warp; pt0 < npatch; pt0 += 4 * (kMixTcThreads / 32)) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        const int c = (pt % (T / 8)) * 8 + c_l, b = (pt / (T / 8)) * 4 + b_l;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt < npatch && c < nc && b < B) {
          const float2* src = src_base + ((long long)b * C + c0 + c) * M + pm0;
          if (TM == 2) {
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(src));
          } else {
            const float2 t = ldg_ordered(src);
            v[u].x = t.x; v[u].y = t.y;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pt = pt0 + u * (kMixTcThreads / 32);
        if (pt < npatch) {
          const int c = (pt % (T / 8)) * 8 + c_l, b = (pt / (T / 8)) * 4 + b_l;
          const float vals[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          for (int t = 0; t < TM; ++t) {
            float h0, l0, h1, l1;
            split_tf32(vals[2 * t], h0, l0);
            split_tf32(vals[2 * t + 1], h1, l1);
            if (is_a) {
              unsigned char* ah = smem + so.a + t * 2 * so.a_one;
              const int o_re = kmajor_off(c, b, 128), o_im = kmajor_off(T + c, b, 128);
              *reinterpret_cast<float*>(ah + o_re) = h0;
              *reinterpret_cast<float*>(ah + so.a_one + o_re) = l0;
              *reinterpret_cast<float*>(ah + o_im) = h1;
              *reinterpret_cast<float*>(ah + so.a_one + o_im) = l1;
            } else {
              unsigned char* bi = smem + so.b + t * so.b_one;
              *reinterpret_cast<float*>(bi + kmajor_off(c, b, NB)) = h0;
              *reinterpret_cast<float*>(bi + kmajor_off(T + c, b, NB)) = h1;
              *reinterpret_cast<float*>(bi + kmajor_off(NH + c, b, NB)) = l0;
              *reinterpret_cast<float*>(bi + kmajor_off(NH + T + c, b, NB)) = l1;
            }
          }
        }
      }
    }
  };
  stage(p.xhat, p.CI, i0, ni, true);
  stage(p.ghat, p.CO, o0, no, false);
  if (p.glb != nullptr && pm0 == 0 && ti == 0 && tid < no) {   // the CTAs holding mode 0: one per output-channel tile
    float acc = 0.f;
    for (int b0 = 0; b0 < B; b0 += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = ldg_ordered(reinterpret_cast<const float*>(p.ghat + ((long long)min(b0 + u, B - 1) * p.CO + o0 + tid) * M));
        if (b0 + u >= B) v[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    p.glb[o0 + tid] = acc * p.s_total;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc_all = make_idesc_tf32(128, NB), idesc_hi = make_idesc_tf32(128, NH);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t lbo_a = 16 * 128, lbo_b = 16 * (uint32_t)NB;
    const uint64_t inc_a = (uint64_t)((2 * lbo_a) >> 4), inc_b = (uint64_t)((2 * lbo_b) >> 4);
    for (int t = 0; t < TM; ++t) {
      uint64_t ah = make_desc(sbase + so.a + t * 2 * so.a_one, lbo_a, 128);
      uint64_t al = make_desc(sbase + so.a + t * 2 * so.a_one + so.a_one, lbo_a, 128);
      uint64_t bb = make_desc(sbase + so.b + t * so.b_one, lbo_b, 128);
      const uint32_t d = tmem + (uint32_t)(t * NB);
      for (int s = 0; s < Kp / 8; ++s) {
        if (leader) {
          mma_tf32(d, ah, bb, idesc_all, s > 0 ? 1u : 0u);
          mma_tf32(d, al, bb, idesc_hi, 1u);
        }
        ah += inc_a; al += inc_a; bb += inc_b;
      }
    }
    if (leader) mma_commit(&bar_mma);
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();

  // ---- epilogue: all 8 warps; warp w reads TMEM lane quadrant w % 4, column half w / 4 (o in [32 wg, 32 wg + 32)) ----
  float* xch = reinterpret_cast<float*>(smem + so.xch);
  const int wq = warp & 3, wg = warp >> 2;
  const int row = wq * 32 + lane;
  const bool is_re = row < ni, is_im = row >= T && row < T + ni;
  int blk, off0;
  decode_mode(p.md, pm0, blk, off0);
  float2* gb = blk == 0 ? p.gw[0] : (blk == 1 ? p.gw[1] : (blk == 2 ? p.gw[2] : p.gw[3]));
  float acc[2][64];                                            // [mode][re(32 o) | im(32 o)]
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    if (t >= TM) break;
    float tr[32], tii[32];                                     // T[row, 32 wg + j] (Gr part), T[row, T + 32 wg + j] (Gi part)
    const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(t * NB + wg * 32);
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 8) {
      float a0[8], a1[8], c0[8], c1[8];
      tmem_ld8(ta + (uint32_t)j0, a0);
      tmem_ld8(ta + (uint32_t)(NH + j0), a1);
      tmem_ld8(ta + (uint32_t)(T + j0), c0);
      tmem_ld8(ta + (uint32_t)(NH + T + j0), c1);
#pragma unroll
      for (int j = 0; j < 8; ++j) { tr[j0 + j] = a0[j] + a1[j]; tii[j0 + j] = c0[j] + c1[j]; }
    }
    if (is_im) {
      float* dst = xch + (row - T) * so.xch_pitch + wg * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) { dst[j] = tr[j]; dst[T + j] = tii[j]; }      // (Xi Gr)[o], (Xi Gi)[o]
    }
    __syncthreads();
    if (is_re) {
      const float* src = xch + row * so.xch_pitch + wg * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        acc[t][j] = tr[j] + src[T + j];                          // Re: Xr Gr + Xi Gi
        acc[t][32 + j] = tii[j] - src[j];                        // Im: Xr Gi - Xi Gr
      }
    }
    __syncthreads();
  }
  if (is_re) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int o = wg * 32 + j;
      if (o < no) {
        float2* dst = gb + ((long long)(i0 + row) * p.CO + o0 + o) * p.md.Mblk + off0;
        if (TM == 2) *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[0][32 + j], acc[1][j], acc[1][32 + j]);
        else *dst = make_float2(acc[0][j], acc[0][32 + j]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace tc
}  // namespace deno
