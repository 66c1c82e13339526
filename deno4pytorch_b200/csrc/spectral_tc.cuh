// spectral_tc.cuh - tcgen05 / TMEM tensor-core kernels of the spectral path (sm_100a only).
//
// All GEMM-shaped work of a stage runs on the 5th-generation tensor cores with fp32-level accuracy
// through the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo; a*b ~ a_hi*b_hi + a_lo*b_hi + a_hi*b_lo,
// relative error ~2^-21; hardware truncates fp32 operands to tf32, so a_hi is the explicitly masked
// value and the products are exact - measured 6e-7 end to end by tools/tc_probe.cu).
//
// Operand convention (pinned on hardware by tools/tc_probe.cu and tools/tc_layout_scan.cu):
// kind::tf32, cta_group::1, M = 128, both operands K-major, SWIZZLE_NONE.  An operand with R rows
// (M rows of A / N rows of B) and K columns is stored as 128-byte core matrices of 8 rows x 4 columns:
//     byte(r, k) = (r % 8) * 16 + (r / 8) * SBO + (k % 4) * 4 + (k / 4) * LBO
// Every image here uses SBO = 128 (row groups contiguous) and LBO = 128 * (R / 8); one MMA consumes
// K = 8, i.e. two core-matrix columns, so consecutive K steps advance the start address by 2 * LBO.
// The accumulator D[m][n] lands in TMEM lane m, column n.
#pragma once
#include "spectral_simt.cuh"

namespace deno {
namespace tc {

// Phase-timing scratch (measurement aid, compiled only with -DDENO_PHASE_PROBE: tools/phase_probe.py builds that
// variant into tools/_build): block 0 of the warp-specialised kernels accumulates clock64() deltas per role and
// phase here; read back with deno_debug_read().  Slots are documented at the writers.  The product build has
// neither the global, nor the clock reads, nor the atomics.
#ifdef DENO_PHASE_PROBE
__device__ long long g_dbg[64];
#define DENO_CLK() clock64()
#define DENO_DBG_ADD(slot, dt) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&g_dbg[slot]), (unsigned long long)(dt)); } while (0)
#else
#define DENO_CLK() 0LL
#define DENO_DBG_ADD(slot, dt) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
// one bulk async copy global -> shared (multiple of 16 bytes, 16-byte aligned both sides); completion on an mbarrier
// that was armed with mbarrier.arrive.expect_tx for at least these bytes
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barriers: producers arrive without blocking, the consumer warp syncs (count = all participating threads)
__device__ __forceinline__ void nbar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// warp-collective
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell); layout type 0 = no swizzle
  return d;
}

// kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// three 16-column TMEM loads (the split-pass tiles of one accumulator, `stride` columns apart) behind ONE wait
__device__ __forceinline__ void tmem_ld16x3_sum(uint32_t taddr, uint32_t stride, float (&v)[16]) {
  uint32_t a[16], b[16], c[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
        "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(taddr + stride));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]),
        "=r"(c[8]), "=r"(c[9]), "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15])
      : "r"(taddr + 2 * stride));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(a[j]) + (__uint_as_float(b[j]) + __uint_as_float(c[j]));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}

using deno::kmajor_off;

// ----------------------------------------------------------------------------------------------
// plane_synthesis_tc: same contract as plane_synthesis_kernel (spectral_simt.cuh).
//   D[y, o] = sum_i x[i, y] * Wl[o, i]                       (bypass 1x1 conv,   K = CIP)
//           + sum_l cos(psi_l y) * U_r[o, l] - sin(psi_l y) * U_i[o, l]   (last-axis synthesis, K = KE)
// One CTA = (batch entry b', R consecutive rows, one 128-wide column chunk); per row one M = 128 tile.
// ----------------------------------------------------------------------------------------------
struct SynthTcParams {
  SynthParams s;          // geometry, pointers, scale, herm (act is a template argument)
  const float* ae;        // pre-split twiddle images [chunk][hi|lo][KE * 128] floats
  int KE;                 // 2 * MWp rounded up to a multiple of 8
  int nchunks;            // column chunks of 128 per row
  int tmem_cols;          // power of two >= 6 * COT (two row buffers x three split-pass tiles)
  const float* urow;      // row-axis synthesis U, pre-split B operand images [NB'][H][cot][hi|lo][KE * COT] (row_synthesis_kernel)
  int COT, nco;           // output channels per CTA (the GEMM N), CO / COT channel tiles (width 128: two tiles of 64)
  int npass;              // K passes of CIP input channels each through the one A_X buffer (width 128: two passes of 64)
};

// ----------------------------------------------------------------------------------------------
// row_synthesis: the row-axis half of the inverse transform, once per (batch entry, row) for the whole layer:
//   U[x][o][l] = f_l * sum_k spec[o][k][l] e^{+i phi(k, x)}
// written as the tf32 hi / lo B operand images plane_synthesis_tc feeds to the last-axis GEMM
// (element (n = o, k = 2l): Re, (o, 2l + 1): -Im), so that kernel's per-tile "stage A" is one bulk copy.
// Doing this inside every synthesis CTA cost half of that kernel (profiles/r2m_phase_probe.txt: 26 k of 58 k
// cycles): each CTA re-read the batch entry's whole spectrum for its 7 rows.
// Block = (batch entry, kRowSynthRows rows); thread = (o, l) pair with all rows in registers.
// ----------------------------------------------------------------------------------------------
constexpr int kRowSynthRows = 8;
struct RowSynthParams {
  const float2* spec;    // [NB'][CO][KH][MW]
  const float2* twH;     // [H][KH4]
  const float* cl;
  float* u;              // [NB'][H][hi|lo][KE * CO]
  float scale;
  int herm;
  int H, KH, KH4, MW, CO, KE;
  int Qp;                // padded per-channel pitch of the staged spectrum (bank-conflict-free (o, l) lanes)
  int tiles;             // ceil(H / kRowSynthRows)
  int ocb, ogroups;      // output channels per block, ceil(CO / ocb); gridDim.x = NB' * tiles * ogroups
  int COT, nco;          // channel tiling of the images: [NB'][H][nco][hi|lo][KE * COT]
};

__host__ __device__ inline int row_synth_pitch(int KH, int MW, int KE) {
  int qp = KH * MW;
  while ((qp - KE / 2) % 16 != 0) ++qp;
  return qp;
}
__host__ __device__ inline size_t row_synth_smem_bytes(int ocb, int KH, int MW, int KE) {
  return ((size_t)ocb * row_synth_pitch(KH, MW, KE) + (size_t)KH * kRowSynthRows) * sizeof(float2);   // spectrum, row twiddles
}

__global__ void __launch_bounds__(384) row_synthesis_kernel(RowSynthParams p) {
  DENO_PDL_SYNC();
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float2* sps = reinterpret_cast<float2*>(rs_smem);                 // [ocb][Qp]
  float2* twr = sps + (size_t)p.ocb * p.Qp;                         // [KH][kRowSynthRows]
  const int tid = threadIdx.x, nthr = blockDim.x;
  int bid = blockIdx.x;
  const int og = bid % p.ogroups; bid /= p.ogroups;
  const int bp = bid / p.tiles;
  const int x0 = (bid - bp * p.tiles) * kRowSynthRows;
  const int trows = min(kRowSynthRows, p.H - x0);
  const int Q = p.KH * p.MW;
  const int o0 = og * p.ocb, noc = min(p.ocb, p.CO - o0);
  const float2* src = p.spec + ((long long)bp * p.CO + o0) * Q;
  {  // The block's channels are one contiguous run of the spectrum.  ALL global loads of the prologue (one row
    // twiddle and up to twelve 16-byte spectrum loads per thread) are issued before the first one is consumed: the
    // kernel is latency-bound and every extra round trip is ~1 us of its ~12.
    const int ntw = p.KH * kRowSynthRows;
    float2 twv = make_float2(0.f, 0.f);
    if (tid < ntw) {
      const int k = tid / kRowSynthRows, rr = tid - k * kRowSynthRows;
      if (rr < trows) twv = ldg_ordered(p.twH + (x0 + rr) * p.KH4 + k);
    }
    const float4* src4 = reinterpret_cast<const float4*>(src);
    const int n4 = noc * Q / 2;                                     // Q = KH * MW is even (KH = 2 * modes)
    for (int i0 = 0; i0 < n4; i0 += 12 * nthr) {
      float4 tmp[12];
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        const int i = min(i0 + u * nthr + tid, n4 - 1);
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(tmp[u].x), "=f"(tmp[u].y), "=f"(tmp[u].z), "=f"(tmp[u].w) : "l"(src4 + i));
      }
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        const int i = i0 + u * nthr + tid;
        if (i < n4) {
          const int ol = (2 * i) / Q, r = 2 * i - ol * Q;
          *reinterpret_cast<float4*>(sps + ol * p.Qp + r) = tmp[u];
        }
      }
    }
    if (tid < ntw) twr[tid] = twv;
    for (int idx = tid + nthr; idx < ntw; idx += nthr) {            // more twiddles than threads (small blocks only)
      const int k = idx / kRowSynthRows, rr = idx - k * kRowSynthRows;
      twr[idx] = rr < trows ? p.twH[(x0 + rr) * p.KH4 + k] : make_float2(0.f, 0.f);
    }
  }
  __syncthreads();
  const int LH = p.KE / 2;                                          // (cos, sin) column pairs incl. zero padding
  const size_t img = (size_t)p.KE * p.COT;                          // floats per hi (or lo) image of one channel tile
  const size_t rowpitch = (size_t)p.nco * 2 * img;                  // floats between consecutive rows
  // (8-channel blocks that stage their results in shared memory and store full 128-byte lines were tried and measured
  // slower - 22 vs 16 us - as were lane pairs splitting the k range)
  for (int task = tid; task < noc * LH; task += nthr) {
    const int ol = task / LH, l = task - ol * LH;
    const int o = o0 + ol;
    float ar[kRowSynthRows], ai[kRowSynthRows];
#pragma unroll
    for (int rr = 0; rr < kRowSynthRows; ++rr) ar[rr] = ai[rr] = 0.f;
    float f = 0.f;
    if (l < p.MW) {
      const float2* sp = sps + ol * p.Qp + l;
#pragma unroll 4
      for (int k = 0; k < p.KH; ++k) {
        const float2 sv = sp[k * p.MW];
        const float4* e4 = reinterpret_cast<const float4*>(twr + k * kRowSynthRows);
#pragma unroll
        for (int r2 = 0; r2 < kRowSynthRows / 2; ++r2) {
          const float4 e = e4[r2];                                  // two rows' twiddles per shared-memory load
          ar[2 * r2] = fmaf(sv.x, e.x, fmaf(-sv.y, e.y, ar[2 * r2]));
          ai[2 * r2] = fmaf(sv.x, e.y, fmaf(sv.y, e.x, ai[2 * r2]));
          ar[2 * r2 + 1] = fmaf(sv.x, e.z, fmaf(-sv.y, e.w, ar[2 * r2 + 1]));
          ai[2 * r2 + 1] = fmaf(sv.x, e.w, fmaf(sv.y, e.z, ai[2 * r2 + 1]));
        }
      }
      f = p.scale * (p.herm ? p.cl[l] : 1.0f);
    }
    const int ot = o / p.COT, oi = o - ot * p.COT;
    const int off = kmajor_off(oi, 2 * l, p.COT) / 4;               // floats; columns 2l and 2l + 1 are adjacent
    float* dst = p.u + ((size_t)bp * p.H + x0) * rowpitch + (size_t)ot * 2 * img + off;
#pragma unroll
    for (int rr = 0; rr < kRowSynthRows; ++rr) {
      if (rr < trows) {
        float h0, l0, h1, l1;
        split_tf32(ar[rr] * f, h0, l0);
        split_tf32(-ai[rr] * f, h1, l1);
        *reinterpret_cast<float2*>(dst + (size_t)rr * rowpitch) = make_float2(h0, h1);
        *reinterpret_cast<float2*>(dst + (size_t)rr * rowpitch + img) = make_float2(l0, l1);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// row_synthesis_tc: the same U images as row_synthesis_kernel, as ONE real GEMM per (batch entry, channel tile,
// column group) on the tensor cores.  With A[x, k] = cos phi(k, x), A[x, KH2 + k] = sin phi(k, x) (constant, host-built
// hi / lo images) and, per image column n = (o, 2l + r),
//     S[n(o,l,0), k] =  f_l Re Y^[o,k,l]    S[n(o,l,0), KH2 + k] = -f_l Im Y^[o,k,l]        ->  D[x, n] =  Re U
//     S[n(o,l,1), k] = -f_l Im Y^[o,k,l]    S[n(o,l,1), KH2 + k] = -f_l Re Y^[o,k,l]        ->  D[x, n] = -Im U
// D[x, n] = sum_k' A[x, k'] S[n, k'] is exactly the pair (Re, -Im) the synthesis kernel's B operand wants, and when the
// N index walks the image in memory order - n = kq * 4 COT + o * 4 + kr for image column 4 kq + kr - the thread that
// owns accumulator row x streams its columns straight into row x's image.  3xTF32 into one accumulator
// (hi*hi + lo*hi + hi*lo).  The SIMT kernel spent 14 us per launch at the Darcy shape on 69 M fp32 FMAs at ~10 % of the
// FMA pipe; here the arithmetic is 18-36 instructions per CTA.
// CTA = (batch entry, channel tile, G column quads); rows in tiles of 128 (ntm <= 2).
// ----------------------------------------------------------------------------------------------
struct RowSynthTcParams {
  const float2* spec;    // [NB'][CO][KH][MW]
  const float* ta;       // row-twiddle A images [ntm][hi|lo][KP * 128] (final tcgen05 layout, from the twiddle table)
  const float* cl;
  float* u;              // [NB'][H][nco][hi|lo][KE * COT]
  float scale;
  int herm;
  int H, KH, MW, CO, KE;
  int KP;                // 2 * round_up(KH, 4): cos columns [0, KP/2), sin columns [KP/2, KP)
  int COT, nco;
  int G, ngrp;           // image column quads per CTA, CTAs per channel tile (G * ngrp = KE / 4)
  int ntm;               // 128-row tiles
};
constexpr int kRsTcThreads = 256;

struct RowSynthTcSmem { int a_one, b_lbo, b_one, a, b, total; };
__host__ __device__ inline RowSynthTcSmem row_synth_tc_smem(int KP, int NT, int ntm) {
  RowSynthTcSmem m;
  m.a_one = 128 * KP * 4;
  m.b_lbo = 16 * NT + 16;                  // padded K-direction stride: lanes walking along k store conflict-free
  m.b_one = (KP / 4) * m.b_lbo;
  m.a = 0;
  m.b = ntm * 2 * m.a_one;
  m.total = m.b + 2 * m.b_one + 2048;      // the M = 128 / N descriptors read whole 8-row groups: keep the tail inside
  const int tiles = (kRsTcThreads / 32) * 32 * (64 + 4) * 4;   // the epilogue's per-warp transpose tiles reuse the operand area
  if (m.total < tiles) m.total = tiles;
  return m;
}

__global__ void __launch_bounds__(kRsTcThreads, 2) row_synthesis_tc_kernel(RowSynthTcParams p) {
  extern __shared__ __align__(128) unsigned char rst_smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_a, bar_mma;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int COT = p.COT, KP = p.KP, KH2 = KP / 2, ntm = p.ntm;
  const int NT = p.G * 4 * COT;
  const RowSynthTcSmem so = row_synth_tc_smem(KP, NT, ntm);
  unsigned char* sa = rst_smem + so.a;
  unsigned char* sb = rst_smem + so.b;
  int bid = blockIdx.x;
  const int grp = bid % p.ngrp; bid /= p.ngrp;
  const int cot = bid % p.nco;
  const int bp = bid / p.nco;
  const int kq0 = grp * p.G;
  uint32_t ncols = 32;
  while ((int)ncols < ntm * NT) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_mma, 1);
    mbar_fence_init();
    mbar_expect_tx(&bar_a, (uint32_t)(ntm * 2 * so.a_one));
    for (int i = 0; i < ntm * 2; ++i)                  // constant images: one bulk copy each
      bulk_g2s(sa + i * so.a_one, p.ta + (size_t)i * (so.a_one / 4), (uint32_t)so.a_one, &bar_a);
  }
  DENO_PDL_SYNC();                                     // the packed spectrum comes from the predecessor (mode mix / outer synthesis)

  // ---- B images: S[n][k'] from the packed spectrum, split hi / lo; lanes walk along k ----
  {
    const int LS = 2 * p.G;                            // last-axis modes of this CTA: l = 2 kq0 + ls
    const int ntask = COT * LS * KH2;
    const float2* src = p.spec + ((long long)bp * p.CO + (long long)cot * COT) * p.KH * p.MW;
    constexpr int NL = 8;                              // loads in flight per thread: the whole slice in one round trip at the usual shapes
    for (int t0 = tid; t0 < ntask; t0 += NL * kRsTcThreads) {
      float2 v[NL];
      int kk[NL], nl[NL];
#pragma unroll
      for (int uu = 0; uu < NL; ++uu) {
        const int t = t0 + uu * kRsTcThreads;
        const int k = t % KH2, q = t / KH2;
        const int ls = q % LS, oi = q / LS;
        const int l = 2 * kq0 + ls;
        kk[uu] = k;
        nl[uu] = (ls >> 1) * 4 * COT + oi * 4 + (ls & 1) * 2;
        v[uu] = make_float2(0.f, 0.f);
        if (t < ntask && k < p.KH && l < p.MW) {
          const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
          const float2 y = ldg_ordered(src + ((long long)oi * p.KH + k) * p.MW + l);
          v[uu] = make_float2(y.x * f, y.y * f);
        }
      }
#pragma unroll
      for (int uu = 0; uu < NL; ++uu) {
        if (t0 + uu * kRsTcThreads < ntask) {
          const int n0 = nl[uu], k = kk[uu];
          const int r0 = (n0 & 7) * 16 + (n0 >> 3) * 128;            // row n0 (r = 0); row n0 + 1 is 16 bytes further
          const int c0 = (k & 3) * 4 + (k >> 2) * so.b_lbo;          // cos column k
          const int c1 = ((KH2 + k) & 3) * 4 + ((KH2 + k) >> 2) * so.b_lbo;   // sin column KH2 + k
          float h, l2;
          split_tf32(v[uu].x, h, l2);                                // +Re: (n0, k);  -Re: (n0 + 1, KH2 + k)
          *reinterpret_cast<float*>(sb + r0 + c0) = h;
          *reinterpret_cast<float*>(sb + so.b_one + r0 + c0) = l2;
          *reinterpret_cast<float*>(sb + r0 + 16 + c1) = -h;
          *reinterpret_cast<float*>(sb + so.b_one + r0 + 16 + c1) = -l2;
          split_tf32(v[uu].y, h, l2);                                // -Im: (n0, KH2 + k) and (n0 + 1, k)
          *reinterpret_cast<float*>(sb + r0 + c1) = -h;
          *reinterpret_cast<float*>(sb + so.b_one + r0 + c1) = -l2;
          *reinterpret_cast<float*>(sb + r0 + 16 + c0) = -h;
          *reinterpret_cast<float*>(sb + so.b_one + r0 + 16 + c0) = -l2;
        }
      }
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    mbar_wait(&bar_a, 0);                              // constant images landed
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_tf32(128, NT);
    const uint32_t sbase = smem_u32(rst_smem);
    const uint32_t lbo_a = 16 * 128, lbo_b = (uint32_t)so.b_lbo;
    const uint64_t inc_a = (uint64_t)((2 * lbo_a) >> 4), inc_b = (uint64_t)((2 * lbo_b) >> 4);
    for (int tm = 0; tm < ntm; ++tm) {
      uint64_t ah = make_desc(sbase + so.a + (2 * tm) * so.a_one, lbo_a, 128);
      uint64_t al = make_desc(sbase + so.a + (2 * tm + 1) * so.a_one, lbo_a, 128);
      uint64_t bh = make_desc(sbase + so.b, lbo_b, 128), bl = make_desc(sbase + so.b + so.b_one, lbo_b, 128);
      const uint32_t d = tmem + (uint32_t)(tm * NT);
      for (int s2 = 0; s2 < KP / 8; ++s2) {
        if (leader) {
          mma_tf32(d, ah, bh, idesc, s2 > 0 ? 1u : 0u);   // hi * hi
          mma_tf32(d, al, bh, idesc, 1u);                 // lo * hi
          mma_tf32(d, ah, bl, idesc, 1u);                 // hi * lo
        }
        ah += inc_a; al += inc_a; bh += inc_b; bl += inc_b;
      }
    }
    if (leader) mma_commit(&bar_mma);
    __syncwarp();
  }
  mbar_wait(&bar_mma, 0);
  fence_after_sync();

  // ---- epilogue: accumulator row x -> tf32 hi / lo -> row x's image (columns in memory order) ----
  // A thread owns one row, so storing straight from registers would put every lane of a store on a different
  // 128-byte line (measured: 31 sectors per request, ~8 k LSU cycles per CTA).  Each warp therefore transposes blocks
  // of CB columns through a private tile in the (now idle) operand area and stores whole row segments.
  {
    const int wq = warp & 3, wh = warp >> 2;           // TMEM lane quadrant of this warp, column half
    const int nhalf = NT / 2;
    const int cbs = (nhalf % 64 == 0) ? 6 : ((nhalf % 32 == 0) ? 5 : 4);   // block width 64 / 32 / 16 columns (nhalf is a multiple of 16)
    const int CB = 1 << cbs;
    const int pitch = CB + 4;                          // floats; 16-byte rows, conflict-free for both access patterns
    float* tile = reinterpret_cast<float*>(rst_smem) + (size_t)warp * 32 * (64 + 4);
    const int q4s = cbs - 2;                           // log2 of the float4 per row of the block
    const int sr = lane >> q4s, sc4 = lane & ((1 << q4s) - 1), srstep = 32 >> q4s;   // this lane's first row / float4 column / row step
    const size_t img = (size_t)p.KE * COT;
    for (int tm = 0; tm < ntm; ++tm) {
      const int xw = tm * 128 + wq * 32;               // first row of this warp
      const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tm * NT + wh * nhalf);
      float* base = p.u + (((size_t)bp * p.H) * p.nco + cot) * 2 * img + (size_t)kq0 * 4 * COT + (size_t)wh * nhalf;
      const size_t rowpitch = (size_t)p.nco * 2 * img;
      for (int cb = 0; cb < nhalf; cb += CB) {
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {         // hi image, then lo image, through the same tile
          for (int c0 = 0; c0 < CB; c0 += 16) {
            float v[16];
            tmem_ld16(ta + (uint32_t)(cb + c0), v);
            float4* dst = reinterpret_cast<float4*>(tile + lane * pitch + c0);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 h, l2;
              split_tf32(v[j], h.x, l2.x);
              split_tf32(v[j + 1], h.y, l2.y);
              split_tf32(v[j + 2], h.z, l2.z);
              split_tf32(v[j + 3], h.w, l2.w);
              dst[j >> 2] = part == 0 ? h : l2;
            }
          }
          __syncwarp();
          {
            const float* src = tile + sr * pitch + sc4 * 4;
            float* dstg = base + (size_t)(xw + sr) * rowpitch + (size_t)part * img + cb + sc4 * 4;
            for (int r = sr; r < 32 && xw + r < p.H; r += srstep) {
              *reinterpret_cast<float4*>(dstg) = *reinterpret_cast<const float4*>(src);
              src += srstep * pitch;
              dstg += (size_t)srstep * rowpitch;
            }
          }
          __syncwarp();
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

struct SynthTcSmem {
  int bw_hi, bw_lo, ae_hi, ae_lo, bu, ax_hi, ax_lo, bias, total;  // byte offsets
  int bu_one;                                                     // bytes of one (hi or lo) U image
};

// CO = output channels of ONE CTA (the channel tile); npass weight images [pass][hi|lo] when the input channels go
// through the A_X buffer in several K passes (bw_lo = offset of pass 0's lo image, pass stride 2 * CIP * CO * 4)
__host__ __device__ inline SynthTcSmem synth_tc_smem(int CIP, int CO, int KE, int R, int npass = 1) {
  SynthTcSmem m;
  int off = 0;
  m.bw_hi = off; off += CIP * CO * 4;
  m.bw_lo = off; off += CIP * CO * 4;
  off += (npass - 1) * 2 * CIP * CO * 4;
  m.ae_hi = off; off += KE * 128 * 4;
  m.ae_lo = off; off += KE * 128 * 4;
  m.bu_one = KE * CO * 4;
  m.bu = off;    off += 2 * R * m.bu_one;   // [row][hi|lo]
  m.ax_hi = off; off += CIP * 128 * 4;
  m.ax_lo = off; off += CIP * 128 * 4;
  m.bias = off;  off += CO * 4;
  m.total = off;
  return m;
}

constexpr int kSynThreads = 256 + 32;   // 8 worker warps + the MMA-issuing warp
constexpr int kSynBarAx = 2;            // A_X images of a row stored (workers arrive, MMA warp syncs)

// Two CTAs per SM up to 32 input channels (96 registers); from 64 channels on the operand images leave room for one CTA
// only, so the register cap is lifted there (at 96 registers the 64-channel instances spilled in the row loop).
// TILED = false (every layer up to 64 channels): one channel tile, one K pass, known at compile time.
// STAGE_A: the row-axis synthesis runs inside this kernel (1-D layers, callers without the U workspace) instead of
// arriving as precomputed images; a template argument so that the usual instances do not carry that code path.
// SQ: the CTA's output-channel tile equals the padded input-channel count (every FNO stack: width in = width out), so the
// GEMM N, the TMEM column offsets and the epilogue's column loops are compile-time constants.
template <int ACT, int CIP, bool TILED, bool STAGE_A, bool SQ>
__global__ void __launch_bounds__(kSynThreads, (CIP <= 32 ? 2 : 1)) plane_synthesis_tc_kernel(SynthTcParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar[2], mbar_u;
  const SynthParams& p = q.s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_mma = warp == 8;             // ninth warp: issues the MMAs, takes no part in the SIMT work
  const int wt = is_mma ? (1 << 28) : tid;     // loop start of the 256 worker threads (beyond every bound for the MMA warp)
  // CO below is the channel tile of THIS CTA (the GEMM N); the layer's output channels are p.CO = q.nco * q.COT
  const int H = p.g.H, W = p.g.W, KH = p.g.KH, MW = p.g.MW, CI = p.g.C, CO = SQ ? CIP : q.COT;
  const int KH4 = round_up(KH, 4);
  const int KE = q.KE, R = p.R;
  const int npass = TILED ? q.npass : 1;
  const int nco = TILED ? q.nco : 1;
  const SynthTcSmem so = synth_tc_smem(CIP, CO, KE, R, npass);
  const int bw_pass = 2 * CIP * CO * 4;            // bytes between the weight images of consecutive K passes
  constexpr int NIT = CIP / 8;            // (point, channel-quad) items per thread: 128 * CIP/4 / 256

  // CTA coordinates
  int bid = blockIdx.x;
  const int chunk = bid % q.nchunks; bid /= q.nchunks;
  int cot = 0;
  if (TILED) { cot = bid % nco; bid /= nco; }     // channel tiles of the same rows on neighbouring CTAs (x re-read from L2)
  const int co0 = cot * CO;
  const int tile = bid % p.tiles;
  const int bp = bid / p.tiles;
  const int x0 = tile * R;
  const int trows = (H > 1) ? min(R, H - x0) : 1;
  const int y0 = chunk * 128;
  const int ny = min(128, W - y0);

  const long long tb0 = DENO_CLK();
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)q.tmem_cols);
  if (tid == 32) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar_u, 1);
    mbar_fence_init();
    // constant twiddle images (final layout in global memory) and, when row_synthesis_kernel ran (q.urow), this
    // tile's U images: bulk async copies, so the prologue costs one memory round trip instead of one per loop trip
    const uint32_t ae_bytes = (uint32_t)(2 * KE * 128 * 4);
    const uint32_t u_bytes = !STAGE_A ? (uint32_t)(trows * 2 * so.bu_one) : 0u;   // one bulk copy per row
    mbar_expect_tx(&mbar_u, ae_bytes + u_bytes);
    bulk_g2s(smem + so.ae_hi, q.ae + (long long)chunk * 2 * KE * 128, ae_bytes, &mbar_u);
  }

  // ---- constant operands -------------------------------------------------------------------
  // bypass weight image B_W[n = o][k = i], split on the fly; four loads in flight per thread
  // (walking the contiguous index of lw in the transposed backward call as well was measured: no change - the 25 %
  // the backward launch loses to the forward one at width >= 64 is not in this prologue)
  for (int ps = 0; ps < npass; ++ps)
    for (int idx0 = wt; idx0 < CO * CIP; idx0 += 4 * 256) {
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = min(idx0 + u * 256, CO * CIP - 1);
        const int o = idx / CIP, i = ps * CIP + (idx - o * CIP);
        wv[u] = (i < CI) ? ldg_ordered(p.lw + (long long)(co0 + o) * p.lw_so + (long long)i * p.lw_si) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * 256;
        if (idx < CO * CIP) {
          const int o = idx / CIP, i = idx - o * CIP;
          float hi, lo;
          split_tf32(wv[u], hi, lo);
          const int off = ps * bw_pass + kmajor_off(o, i, CO);
          *reinterpret_cast<float*>(smem + so.bw_hi + off) = hi;
          *reinterpret_cast<float*>(smem + so.bw_lo + off) = lo;
        }
      }
    }
  float* bias_s = reinterpret_cast<float*>(smem + so.bias);
  for (int o = wt; o < CO; o += 256) bias_s[o] = (p.bias != nullptr) ? p.bias[co0 + o] : 0.f;

  // Everything above depends on constants and weights only.  The U images come from the immediate predecessor
  // (row_synthesis_kernel) and x / gz from earlier kernels: from here on the predecessor must have completed.
  DENO_PDL_SYNC();
  if (!STAGE_A && tid == 32) {
    const size_t img2 = 2 * (size_t)(so.bu_one / 4);                 // floats of one row's hi + lo images of one channel tile
    for (int rr = 0; rr < trows; ++rr)
      bulk_g2s(smem + so.bu + rr * 2 * so.bu_one, q.urow + (((size_t)bp * H + x0 + rr) * nco + cot) * img2,
               (uint32_t)(2 * so.bu_one), &mbar_u);
  }

  // ---- x row prefetch (registers) ------------------------------------------------------------
  float xr[NIT][4];
  const long long ubase = plane_base(p.g, bp, 0);
  auto prefetch = [&](int rr, int ps) {              // row of the tile, K pass
    const long long rowoff = ubase + (long long)(x0 + rr) * W + y0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int item = it * 256 + tid;
      const int pt = item & 127, cq = item >> 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = ps * CIP + cq * 4 + j;
        xr[it][j] = (pt < ny && ch < CI) ? p.u[rowoff + (long long)ch * p.g.sc + pt] : 0.f;
      }
    }
  };
  if (!is_mma) prefetch(0, 0);

  const long long tb1 = DENO_CLK();
  if (warp == 1) DENO_DBG_ADD(48, tb1 - tb0);          // slot 48: TMEM alloc, weight / twiddle images, first prefetch issue
  // ---- stage A: U[r][o][l] = f_l * sum_k spec[o][k][l] e^{+i phi(k, x0 + r)}, written as B operand images ----
  // Normally precomputed (q.urow, bulk copy issued above); the in-kernel version below is kept for callers without
  // the workspace: the packed spectrum of this batch entry is staged through the (still unused) A_X area in chunks
  // of OC output channels with fully coalesced loads, so the k-loop runs out of shared memory.
  if (STAGE_A) {
    unsigned char* bu = smem + so.bu;
    for (int idx = wt; idx < 2 * R * so.bu_one / 16; idx += 256) reinterpret_cast<float4*>(bu)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    float2* twr = reinterpret_cast<float2*>(smem + so.ax_hi);            // [trows][KH] row twiddles
    float2* sps = twr + round_up(kSynthMaxRows * KH, 2);                 // [OC][KH*MW] spectrum chunk
    const int Q = KH * MW;
    const int avail = 2 * CIP * 128 * 4 - round_up(kSynthMaxRows * KH, 2) * 8;
    int OC = avail / (Q * 8);
    if (OC > CO) OC = CO;
    for (int idx = wt; idx < trows * KH; idx += 256) {
      const int rr = idx / KH, k = idx - rr * KH;
      twr[idx] = p.twH[(x0 + rr) * KH4 + k];
    }
    if (warp == 1) DENO_DBG_ADD(57, DENO_CLK() - tb1);   // slot 57: zero fill + row twiddles
    for (int o0 = 0; o0 < CO; o0 += OC) {
      const int oc = min(OC, CO - o0);
      const long long tc0 = DENO_CLK();
      __syncthreads();   // previous chunk consumed (first pass: zero fill / twiddles ordered before use)
      const float2* src = p.spec + ((long long)bp * CO + o0) * Q;
      for (int idx0 = 0; idx0 < oc * Q; idx0 += 8 * 256) {     // eight independent loads in flight per thread
        float2 tmp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) tmp[u] = ldg_ordered(src + min(idx0 + u * 256 + (tid & 255), oc * Q - 1));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = idx0 + u * 256 + wt;
          if (idx < oc * Q) sps[idx] = tmp[u];
        }
      }
      __syncthreads();
      const long long tc1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(58, tc1 - tc0);       // slot 58: spectrum chunk load (incl. barriers)
      // task = (row half, output channel, mode): two threads share an (o, l) pair, four rows each
      for (int task = wt; task < 2 * oc * MW; task += 256) {
        const int half = task / (oc * MW), t2 = task - half * (oc * MW);
        const int ol = t2 / MW, l = t2 - ol * MW;
        const int o = o0 + ol;
        const int r0 = half * (kSynthMaxRows / 2);
        float ar[kSynthMaxRows / 2], ai[kSynthMaxRows / 2];
#pragma unroll
        for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) ar[rr] = ai[rr] = 0.f;
        const float2* sp = sps + ol * Q + l;
#pragma unroll 4
        for (int k = 0; k < KH; ++k) {
          const float2 sv = sp[k * MW];
#pragma unroll
          for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) {
            if (r0 + rr < trows) {
              const float2 e = twr[(r0 + rr) * KH + k];
              ar[rr] = fmaf(sv.x, e.x, fmaf(-sv.y, e.y, ar[rr]));
              ai[rr] = fmaf(sv.x, e.y, fmaf(sv.y, e.x, ai[rr]));
            }
          }
        }
        const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
        const int off = kmajor_off(o, 2 * l, CO);   // k = 2l (cos partner) and 2l+1 (sin partner) share a 16-byte row
#pragma unroll
        for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) {
          if (r0 + rr < trows) {
            float h0, l0, h1, l1;
            split_tf32(ar[rr] * f, h0, l0);
            split_tf32(-ai[rr] * f, h1, l1);
            unsigned char* base = bu + (2 * (r0 + rr)) * so.bu_one + off;
            *reinterpret_cast<float2*>(base) = make_float2(h0, h1);
            *reinterpret_cast<float2*>(base + so.bu_one) = make_float2(l0, l1);
          }
        }
      }
      if (warp == 1) { DENO_DBG_ADD(59, DENO_CLK() - tc1); DENO_DBG_ADD(60, 1); }   // slots 59/60: chunk compute, chunks
    }
    __syncthreads();     // the A_X area is about to be overwritten by the first row tile
  }

  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  mbar_wait(&mbar_u, 0);                               // twiddle (and U) images have landed (barrier initialised before the sync above)
  const long long tb2 = DENO_CLK();
  if (warp == 1) DENO_DBG_ADD(49, tb2 - tb1);          // slot 49: stage A
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, CO);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t lbo_m = 16 * 128;   // images with 128 rows (A operands)
  const uint32_t lbo_n = 16 * CO;    // images with CO rows (B operands)

  const int wq = warp & 3, wg = warp >> 2;          // TMEM lane quadrant / column half
  const int ncol_half = CO / 2;
  const int yl = wq * 32 + lane;                    // this thread's point inside the chunk
  const long long obase_b = (long long)(bp / p.g.NBI) * p.sbo_out + (long long)(bp % p.g.NBI) * p.g.sbi + (long long)co0 * p.g.sc;

  auto epilogue = [&](int rr) {
    const uint32_t tb = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)((rr & 1) * 3 * CO + wg * ncol_half);
    const long long pbase = obase_b + (long long)(x0 + rr) * W + y0 + yl;
    for (int c0 = 0; c0 < ncol_half; c0 += 16) {
      float v[16], v1[16], v2[16];
      tmem_ld16(tb + c0, v);
      tmem_ld16(tb + CO + c0, v1);
      tmem_ld16(tb + 2 * CO + c0, v2);
      if (yl < ny) {
        if (ACT == ACT_GELU) {                           // two channels per packed fp32x2 GELU evaluation
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const int o = wg * ncol_half + c0 + j;
            const float2 bb = *reinterpret_cast<const float2*>(bias_s + o);
            const float2 zz = __fadd2_rn(__fadd2_rn(make_float2(v[j], v[j + 1]), bb),
                                         __fadd2_rn(make_float2(v1[j], v1[j + 1]), make_float2(v2[j], v2[j + 1])));
            float2 yv, dv;
            gelu_fast2(zz, yv, dv);
            const long long a = pbase + (long long)o * p.g.sc;
            if (p.z != nullptr) {
              p.z[a] = dv.x;
              p.z[a + p.g.sc] = dv.y;
            }
            p.y[a] = yv.x;
            p.y[a + p.g.sc] = yv.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int o = wg * ncol_half + c0 + j;
            const float zz = (v[j] + bias_s[o]) + (v1[j] + v2[j]);
            const long long a = pbase + (long long)o * p.g.sc;
            if (p.z != nullptr) {
              float yv, dv;
              act_apply_grad(ACT, zz, yv, dv);
              p.z[a] = dv;
              p.y[a] = yv;
            } else {
              p.y[a] = act_apply(ACT, zz);
            }
          }
        }
      }
    }
  };

  if (is_mma) {
    // ---- MMA warp: per row, wait until the 256 workers have stored the A_X images, issue, commit ----
    const bool leader = elect_one();   // warp-uniform loop (descriptors in uniform registers); one lane issues
    const uint64_t inc_m = (uint64_t)((2 * lbo_m) >> 4), inc_n = (uint64_t)((2 * lbo_n) >> 4);
    for (int step = 0, rr = 0, ps = 0; rr < trows; ++step, ps = (ps + 1 == npass) ? 0 : ps + 1, rr += (ps == 0)) {
      nbar_sync(kSynBarAx, kSynThreads);
      fence_after_sync();
      const long long tr3 = DENO_CLK();
      // Back-to-back MMAs into ONE accumulator serialise on the tensor pipe's latency (measured ~120 cycles each at
      // N = 32), so the three split passes (hi*hi, lo*hi, hi*lo) accumulate into three separate TMEM tiles and are
      // issued round-robin; the epilogue adds the tiles.  Descriptors are built once and advanced by a constant.
      const uint32_t d = tmem + (uint32_t)((rr & 1) * 3 * CO);
      uint64_t axh = make_desc(sbase + so.ax_hi, lbo_m, 128), axl = make_desc(sbase + so.ax_lo, lbo_m, 128);
      uint64_t bwh = make_desc(sbase + so.bw_hi + ps * bw_pass, lbo_n, 128), bwl = make_desc(sbase + so.bw_lo + ps * bw_pass, lbo_n, 128);
#pragma unroll
      for (int s = 0; s < CIP / 8; ++s) {
        const uint32_t acc = (s > 0 || ps > 0) ? 1u : 0u;
        if (leader) {
          mma_tf32(d, axh, bwh, idesc, acc);            // hi * hi
          mma_tf32(d + CO, axl, bwh, idesc, acc);       // lo * hi
          mma_tf32(d + 2 * CO, axh, bwl, idesc, acc);   // hi * lo
        }
        axh += inc_m; axl += inc_m; bwh += inc_n; bwl += inc_n;
      }
      if (ps == npass - 1) {                            // last K pass of the row: the last-axis synthesis on top
        uint64_t aeh = make_desc(sbase + so.ae_hi, lbo_m, 128), ael = make_desc(sbase + so.ae_lo, lbo_m, 128);
        uint64_t buh = make_desc(sbase + so.bu + (2 * rr) * so.bu_one, lbo_n, 128);
        uint64_t bul = make_desc(sbase + so.bu + (2 * rr + 1) * so.bu_one, lbo_n, 128);
        for (int s = 0; s < KE / 8; ++s) {
          if (leader) {
            mma_tf32(d, aeh, buh, idesc, 1u);
            mma_tf32(d + CO, ael, buh, idesc, 1u);
            mma_tf32(d + 2 * CO, aeh, bul, idesc, 1u);
          }
          aeh += inc_m; ael += inc_m; buh += inc_n; bul += inc_n;
        }
      }
      if (leader) mma_commit(&mbar[step & 1]);
      __syncwarp();
      if (leader) DENO_DBG_ADD(53, DENO_CLK() - tr3);   // slot 53: MMA issue
    }
  } else {
    // ---- workers: A_X images of row rr, then (while the tensor core works on it) the epilogue of row rr - 1 ----
    for (int step = 0, rr = 0, ps = 0; rr < trows; ++step, ps = (ps + 1 == npass) ? 0 : ps + 1, rr += (ps == 0)) {
      const long long tr0 = DENO_CLK();
      if (step > 0) mbar_wait(&mbar[(step - 1) & 1], ((step - 1) >> 1) & 1);   // MMAs of the previous step done: A_X free (ps == 0: D(rr-1) ready)
      const long long tr1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(50, tr1 - tr0);        // slot 50: wait for the previous row's MMAs
      // registers -> A_X images (hi / lo)
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int item = it * 256 + tid;
        const int pt = item & 127, cq = item >> 7;
        float4 hi, lo;
        split_tf32(xr[it][0], hi.x, lo.x);
        split_tf32(xr[it][1], hi.y, lo.y);
        split_tf32(xr[it][2], hi.z, lo.z);
        split_tf32(xr[it][3], hi.w, lo.w);
        const int off = kmajor_off(pt, cq * 4, 128);
        *reinterpret_cast<float4*>(smem + so.ax_hi + off) = hi;
        *reinterpret_cast<float4*>(smem + so.ax_lo + off) = lo;
      }
      fence_proxy_async();
      fence_before_sync();                               // also orders the D reads of epilogue(rr - 2) before the hand-off
      nbar_arrive(kSynBarAx, kSynThreads);
      const long long tr4 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(51, tr4 - tr1);        // slot 51: split + store + hand-off
      {
        const int nps = (ps + 1 == npass) ? 0 : ps + 1, nrr = rr + (nps == 0);
        if (nrr < trows) prefetch(nrr, nps);
      }
      const long long tr5 = DENO_CLK();
      if (rr > 0 && ps == 0) {
        fence_after_sync();
        epilogue(rr - 1);
      }
      if (warp == 1) { DENO_DBG_ADD(54, tr5 - tr4); DENO_DBG_ADD(55, DENO_CLK() - tr5); DENO_DBG_ADD(56, 1); }  // prefetch issue, epilogue, rows
    }
  }
  if (!is_mma) {
    const int last = trows * npass - 1;
    mbar_wait(&mbar[last & 1], (last >> 1) & 1);
    fence_after_sync();
    epilogue(trows - 1);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)q.tmem_cols);
}

#ifdef DENO_WITH_SYNTH_WS
// ----------------------------------------------------------------------------------------------
// plane_synthesis_ws: warp-specialised version of plane_synthesis_tc (same math, same operand images).
//   warps 0-15  epilogue : TMEM -> registers -> bias + activation (+ derivative) -> coalesced global stores
//   warps 16-23 loaders  : global x row -> registers (two rows in flight) -> tf32 split -> A_X image (two stages)
//   warp  24    MMA      : one elected lane issues the 3 x (CIP/8 + KE/8) tcgen05.mma of a row
// The roles hand rows to each other through mbarriers (A_X full/empty, D full/empty), so loading row r+1,
// the tensor-core work of row r and the epilogue of row r-1 overlap inside one persistent CTA per SM.
// ----------------------------------------------------------------------------------------------
constexpr int kWsEpiWarps = 16, kWsLoadWarps = 8;   // the epilogue is latency-bound per warp: give it many warps
constexpr int kWsThreads = (kWsEpiWarps + kWsLoadWarps + 1) * 32;

struct SynthWsSmem {
  int bw_hi, bw_lo, ae_hi, ae_lo, bu, ax, bias, total;  // byte offsets
  int bu_one, ax_one;
};

__host__ __device__ inline SynthWsSmem synth_ws_smem(int CIP, int CO, int KE, int R) {
  SynthWsSmem m;
  int off = 0;
  m.bw_hi = off; off += CIP * CO * 4;
  m.bw_lo = off; off += CIP * CO * 4;
  m.ae_hi = off; off += KE * 128 * 4;
  m.ae_lo = off; off += KE * 128 * 4;
  m.bu_one = KE * CO * 4;
  m.bu = off;    off += 2 * R * m.bu_one;        // [row][hi|lo]
  m.ax_one = CIP * 128 * 4;
  m.ax = off;    off += 4 * m.ax_one;            // [stage][hi|lo]
  m.bias = off;  off += CO * 4;
  m.total = off;
  return m;
}

template <int ACT, int CIP>
__global__ void __launch_bounds__(kWsThreads, 1) plane_synthesis_ws_kernel(SynthTcParams q) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_ax_full[2], bar_ax_empty[2], bar_d_full[2], bar_d_empty[2];
  const SynthParams& p = q.s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nt = kWsThreads;
  const int H = p.g.H, W = p.g.W, KH = p.g.KH, MW = p.g.MW, CI = p.g.C, CO = p.CO;
  const int KH4 = round_up(KH, 4);
  const int KE = q.KE, R = p.R;
  const SynthWsSmem so = synth_ws_smem(CIP, CO, KE, R);
  constexpr int NCQ = CIP / 4;            // channel quads per point (one loader thread = one point)

  DENO_PDL_SYNC();
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)q.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_ax_full[s], kWsLoadWarps * 32);
      mbar_init(&bar_ax_empty[s], 1);
      mbar_init(&bar_d_full[s], 1);
      mbar_init(&bar_d_empty[s], kWsEpiWarps * 32);
    }
    mbar_fence_init();
  }
  // bypass weight image B_W[n = o][k = i], split on the fly (constant for the CTA's lifetime)
  for (int idx = tid; idx < CO * CIP; idx += nt) {
    const int o = idx / CIP, i = idx - o * CIP;
    const float w = (i < CI) ? p.lw[(long long)o * p.lw_so + (long long)i * p.lw_si] : 0.f;
    float hi, lo;
    split_tf32(w, hi, lo);
    const int off = kmajor_off(o, i, CO);
    *reinterpret_cast<float*>(smem + so.bw_hi + off) = hi;
    *reinterpret_cast<float*>(smem + so.bw_lo + off) = lo;
  }
  float* bias_s = reinterpret_cast<float*>(smem + so.bias);
  for (int o = tid; o < CO; o += nt) bias_s[o] = (p.bias != nullptr) ? p.bias[o] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_tf32(128, CO);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t lbo_m = 16 * 128, lbo_n = 16 * CO;
  const int nunits = p.g.NBO * p.g.NBI * p.tiles * q.nchunks;
  int cur_chunk = -1;
  int it0 = 0;                            // rows handled so far by this CTA (drives stage / parity of the barriers)

  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    int bid = unit;
    const int chunk = bid % q.nchunks; bid /= q.nchunks;
    const int tile = bid % p.tiles;
    const int bp = bid / p.tiles;
    const int x0 = tile * R;
    const int trows = (H > 1) ? min(R, H - x0) : 1;
    const int y0 = chunk * 128;
    const int ny = min(128, W - y0);

    const long long t_unit0 = DENO_CLK();
    // ---- per-unit setup by all warps: twiddle images (if the column chunk changed) and stage A ----
    if (chunk != cur_chunk) {
      const float4* src = reinterpret_cast<const float4*>(q.ae + (long long)chunk * 2 * KE * 128);
      float4* dst = reinterpret_cast<float4*>(smem + so.ae_hi);
      for (int idx = tid; idx < 2 * KE * 128 / 4; idx += nt) dst[idx] = src[idx];
      cur_chunk = chunk;
    }
    {
      unsigned char* bu = smem + so.bu;
      for (int idx = tid; idx < 2 * R * so.bu_one / 16; idx += nt) reinterpret_cast<float4*>(bu)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      float2* twr = reinterpret_cast<float2*>(smem + so.ax);               // [trows][KH] row twiddles
      float2* sps = twr + round_up(kSynthMaxRows * KH, 2);                 // [OC][KH*MW] spectrum chunk
      const int Q = KH * MW;
      const int avail = 4 * so.ax_one - round_up(kSynthMaxRows * KH, 2) * 8;
      int OC = avail / (Q * 8);
      if (OC > CO) OC = CO;
      for (int idx = tid; idx < trows * KH; idx += nt) {
        const int rr = idx / KH, k = idx - rr * KH;
        twr[idx] = p.twH[(x0 + rr) * KH4 + k];
      }
      for (int o0 = 0; o0 < CO; o0 += OC) {
        const int oc = min(OC, CO - o0);
        __syncthreads();
        const float2* src = p.spec + ((long long)bp * CO + o0) * Q;
        for (int idx0 = 0; idx0 < oc * Q; idx0 += 8 * nt) {     // eight independent loads in flight per thread
          float2 tmp[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * nt + tid;
            tmp[u] = ldg_ordered(src + min(idx, oc * Q - 1));
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * nt + tid;
            if (idx < oc * Q) sps[idx] = tmp[u];
          }
        }
        __syncthreads();
        // task = (row half, output channel, mode): two threads share an (o, l) pair, four rows each
        for (int task = tid; task < 2 * oc * MW; task += nt) {
          const int half = task / (oc * MW), t2 = task - half * (oc * MW);
          const int ol = t2 / MW, l = t2 - ol * MW;
          const int o = o0 + ol;
          const int r0 = half * (kSynthMaxRows / 2);
          float ar[kSynthMaxRows / 2], ai[kSynthMaxRows / 2];
#pragma unroll
          for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) ar[rr] = ai[rr] = 0.f;
          const float2* sp = sps + ol * Q + l;
#pragma unroll 4
          for (int k = 0; k < KH; ++k) {
            const float2 sv = sp[k * MW];
#pragma unroll
            for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) {
              if (r0 + rr < trows) {
                const float2 e = twr[(r0 + rr) * KH + k];
                ar[rr] = fmaf(sv.x, e.x, fmaf(-sv.y, e.y, ar[rr]));
                ai[rr] = fmaf(sv.x, e.y, fmaf(sv.y, e.x, ai[rr]));
              }
            }
          }
          const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
          const int off = kmajor_off(o, 2 * l, CO);
#pragma unroll
          for (int rr = 0; rr < kSynthMaxRows / 2; ++rr) {
            if (r0 + rr < trows) {
              float h0, l0, h1, l1;
              split_tf32(ar[rr] * f, h0, l0);
              split_tf32(-ai[rr] * f, h1, l1);
              unsigned char* base = bu + (2 * (r0 + rr)) * so.bu_one + off;
              *reinterpret_cast<float2*>(base) = make_float2(h0, h1);
              *reinterpret_cast<float2*>(base + so.bu_one) = make_float2(l0, l1);
            }
          }
        }
      }
      fence_proxy_async();
      __syncthreads();     // operand images complete; the A_X stages are free for the loaders
    }
    if (warp == 0) DENO_DBG_ADD(0, DENO_CLK() - t_unit0);                 // slot 0: unit setup (stage A)
    const long long t_roles0 = DENO_CLK();

    if (warp < kWsEpiWarps) {
      // ===== epilogue warps: lane quadrant wq of TMEM, column slice wg =====
      const int wq = warp & 3, wg = warp >> 2;
      const int ncol_per = CO / (kWsEpiWarps / 4);          // columns per warp (multiple of 8)
      const int yl = wq * 32 + lane;
      const long long obase_b = (long long)(bp / p.g.NBI) * p.sbo_out + (long long)(bp % p.g.NBI) * p.g.sbi;
      for (int rr = 0; rr < trows; ++rr) {
        const int i = it0 + rr, st = i & 1;
        const long long te0 = DENO_CLK();
        mbar_wait(&bar_d_full[st], (uint32_t)((i >> 1) & 1));
        fence_after_sync();
        const long long te1 = DENO_CLK();
        if (warp == 0) DENO_DBG_ADD(1, te1 - te0);                       // slot 1: epilogue waits for D
        const uint32_t tb = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(st * 3 * CO + wg * ncol_per);
        const long long pbase = obase_b + (long long)(x0 + rr) * W + y0 + yl;
        for (int c0 = 0; c0 < ncol_per; c0 += 8) {
          float v[8], v1[8], v2[8];
          tmem_ld8(tb + c0, v);
          tmem_ld8(tb + CO + c0, v1);
          tmem_ld8(tb + 2 * CO + c0, v2);
          if (c0 + 8 >= ncol_per) {            // last TMEM read of this row by this thread: release the buffer early
            fence_before_sync();
            mbar_arrive(&bar_d_empty[st]);
          }
          if (yl < ny) {
            float zz[8], yv[8], dv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) zz[j] = (v[j] + bias_s[wg * ncol_per + c0 + j]) + (v1[j] + v2[j]);
            if (p.z != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) act_apply_grad(ACT, zz[j], yv[j], dv[j]);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const long long a = pbase + (long long)(wg * ncol_per + c0 + j) * p.g.sc;
                p.z[a] = dv[j];
                p.y[a] = yv[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) p.y[pbase + (long long)(wg * ncol_per + c0 + j) * p.g.sc] = act_apply(ACT, zz[j]);
            }
          }
        }
        if (warp == 0) DENO_DBG_ADD(2, DENO_CLK() - te1);                 // slot 2: epilogue work per row
      }
    } else if (warp < kWsEpiWarps + kWsLoadWarps) {
      // ===== loader warps: one thread = one point of the row and half of the channel quads; two rows of
      // loads are kept in flight (register double buffer) so the DRAM latency never gates the MMA warp =====
      const int lt = tid - kWsEpiWarps * 32;
      const int pt = lt & 127, chalf = lt >> 7;
      constexpr int NQ = NCQ / 2 > 0 ? NCQ / 2 : 1;        // channel quads per loader thread
      const int cq0 = chalf * NQ;
      const long long ubase = plane_base(p.g, bp, 0) + y0 + pt;
      float xa[NQ][4], xb[NQ][4];
      auto prefetch = [&](int rr, float (&xr)[NQ][4]) {
        const long long rowoff = ubase + (long long)(x0 + rr) * W;
#pragma unroll
        for (int c = 0; c < NQ; ++c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = (cq0 + c) * 4 + j;
            xr[c][j] = (pt < ny && ch < CI && cq0 + c < NCQ) ? p.u[rowoff + (long long)ch * p.g.sc] : 0.f;
          }
        }
      };
      auto publish = [&](int rr, float (&xr)[NQ][4]) {
        const int i = it0 + rr, st = i & 1;
        const long long tl0 = DENO_CLK();
        mbar_wait(&bar_ax_empty[st], (uint32_t)(((i >> 1) & 1) ^ 1));
        const long long tl1 = DENO_CLK();
        if (warp == kWsEpiWarps) DENO_DBG_ADD(3, tl1 - tl0);             // slot 3: loader waits for a free stage
        unsigned char* axh = smem + so.ax + st * 2 * so.ax_one;
#pragma unroll
        for (int c = 0; c < NQ; ++c) {
          if (cq0 + c < NCQ) {
            float4 hi, lo;
            split_tf32(xr[c][0], hi.x, lo.x);
            split_tf32(xr[c][1], hi.y, lo.y);
            split_tf32(xr[c][2], hi.z, lo.z);
            split_tf32(xr[c][3], hi.w, lo.w);
            const int off = kmajor_off(pt, (cq0 + c) * 4, 128);
            *reinterpret_cast<float4*>(axh + off) = hi;
            *reinterpret_cast<float4*>(axh + so.ax_one + off) = lo;
          }
        }
        fence_proxy_async();
        mbar_arrive(&bar_ax_full[st]);
        if (warp == kWsEpiWarps) DENO_DBG_ADD(4, DENO_CLK() - tl1);       // slot 4: loader split + store (incl. wait for LDG data)
      };
      prefetch(0, xa);
      if (trows > 1) prefetch(1, xb);
      for (int rr = 0; rr < trows; rr += 2) {
        publish(rr, xa);
        if (rr + 2 < trows) prefetch(rr + 2, xa);
        if (rr + 1 < trows) {
          publish(rr + 1, xb);
          if (rr + 3 < trows) prefetch(rr + 3, xb);
        }
      }
    } else {
      // ===== MMA warp =====
      const bool leader = elect_one();
      const uint64_t inc_m = (uint64_t)((2 * lbo_m) >> 4), inc_n = (uint64_t)((2 * lbo_n) >> 4);
      for (int rr = 0; rr < trows; ++rr) {
        const int i = it0 + rr, st = i & 1;
        const uint32_t par = (uint32_t)((i >> 1) & 1);
        const long long tm0 = DENO_CLK();
        mbar_wait(&bar_ax_full[st], par);
        const long long tm1 = DENO_CLK();
        mbar_wait(&bar_d_empty[st], par ^ 1);
        fence_after_sync();
        const long long tm2 = DENO_CLK();
        DENO_DBG_ADD(5, tm1 - tm0);                                      // slot 5: MMA warp waits for A_X
        DENO_DBG_ADD(6, tm2 - tm1);                                      // slot 6: MMA warp waits for a free accumulator
        const uint32_t d = tmem + (uint32_t)(st * 3 * CO);
        const uint32_t axs = sbase + so.ax + st * 2 * so.ax_one;
        uint64_t axh = make_desc(axs, lbo_m, 128), axl = make_desc(axs + so.ax_one, lbo_m, 128);
        uint64_t bwh = make_desc(sbase + so.bw_hi, lbo_n, 128), bwl = make_desc(sbase + so.bw_lo, lbo_n, 128);
#pragma unroll
        for (int s = 0; s < CIP / 8; ++s) {
          const uint32_t acc = s > 0 ? 1u : 0u;
          if (leader) {
            mma_tf32(d, axh, bwh, idesc, acc);            // hi * hi
            mma_tf32(d + CO, axl, bwh, idesc, acc);       // lo * hi
            mma_tf32(d + 2 * CO, axh, bwl, idesc, acc);   // hi * lo
          }
          axh += inc_m; axl += inc_m; bwh += inc_n; bwl += inc_n;
        }
        uint64_t aeh = make_desc(sbase + so.ae_hi, lbo_m, 128), ael = make_desc(sbase + so.ae_lo, lbo_m, 128);
        uint64_t buh = make_desc(sbase + so.bu + (2 * rr) * so.bu_one, lbo_n, 128);
        uint64_t bul = make_desc(sbase + so.bu + (2 * rr + 1) * so.bu_one, lbo_n, 128);
        for (int s = 0; s < KE / 8; ++s) {
          if (leader) {
            mma_tf32(d, aeh, buh, idesc, 1u);
            mma_tf32(d + CO, ael, buh, idesc, 1u);
            mma_tf32(d + 2 * CO, aeh, bul, idesc, 1u);
          }
          aeh += inc_m; ael += inc_m; buh += inc_n; bul += inc_n;
        }
        if (leader) {
          mma_commit(&bar_ax_empty[st]);     // the A_X stage may be overwritten once these MMAs retire
          mma_commit(&bar_d_full[st]);       // ... and the accumulators are ready for the epilogue warps
        }
        __syncwarp();
        DENO_DBG_ADD(7, DENO_CLK() - tm2);                                // slot 7: MMA issue
      }
    }
    if (warp == 0) { DENO_DBG_ADD(8, DENO_CLK() - t_roles0); DENO_DBG_ADD(9, trows); }   // slots 8/9: role-phase cycles, rows
    it0 += trows;
    fence_before_sync();
    __syncthreads();      // unit drained: every role is done with the images, the A_X stages and TMEM
    fence_after_sync();
  }
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)q.tmem_cols);
}

#endif  // DENO_WITH_SYNTH_WS

// ----------------------------------------------------------------------------------------------
// plane_analysis_tc: same contract as plane_analysis_kernel (spectral_simt.cuh) for planes with
// H <= 128 rows.  Two chained GEMMs per plane, both on the tensor cores:
//   stage 1  D1[x, n]  = sum_y u[x, y] * E1[n, y]          n = 2l: cos, 2l+1: -sin      (K = Wp)
//            -> D1[x, 2l] + i D1[x, 2l+1] = T[x, l] = sum_y u[x,y] e^{-i psi_l y}
//   stage 2  D2[r, n]  = sum_x A2[r, x] * D1[x, n]          (K = Hp), rows r = k: cos(phi_k x), r = KS0 + k: sin(phi_k x)
//            -> X^[k, l] = (D2[k,2l] + D2[KS0+k,2l+1]) + i (D2[k,2l+1] - D2[KS0+k,2l])     (e^{-i phi} (Tr + i Ti))
// 3xTF32: the B operand of each stage is ONE image with the hi rows followed by the lo rows, so
// A_hi x [B_hi | B_lo] is a single N = 2 N1 instruction and A_lo x B_hi a second one (an instruction costs
// the same for every N <= 128, tools/tc_mma_bench.cu).
// One persistent CTA per SM; 16 worker warps + 1 MMA-issuing warp, software-pipelined over planes:
//   workers : wait stage 1(q) | D1(q) TMEM -> split -> B2 image | plane q+1: (gy * act') -> A images |
//             wait stage 2(q) | combine cos / sin rows of D2(q) through shared memory -> global
//   MMA warp: stage 2(q) as soon as B2 is stored | bulk copy of plane q+2, L2 prefetch of q+3 | stage 1(q+1)
// so the tensor core works on plane q+1 while the workers finish plane q.  Hand-offs to the MMA warp are
// bar.arrive / bar.sync named barriers; completions come back through mbarriers (tcgen05.commit).
// ----------------------------------------------------------------------------------------------
struct AnalysisTcParams {
  PlaneGeom g;
  const float* u;        // x, or gy when FUSE_GZ
  const float* z;        // FUSE_GZ: saved act'(z)
  float* gz_out;
  float2* out;           // [nplanes][KH][MW]
  const float* e1;       // stacked [hi rows | lo rows] image, 2 N1 rows x Wp
  const float* a2;       // [hi|lo] images, KRp rows x Hp
  const float* cl;
  float scale;
  int herm;
  int act;
  int nplanes;
  int N1, Wp, Hp, KS0, KRp;
  int nsub, HS;          // planes taller than 128 rows: nsub row slabs of HS rows (Hp = HS padded to 8), stage 2 accumulates over them
  // Staging ring (chosen by the launcher from what fits next to the operand images): `nslots` slots of one slab each,
  // filled by bulk copies issued nslots items ahead; with `zstage` a slot also holds the slab's act'(z) (FUSE_GZ), so the
  // gz pass reads both factors from shared memory instead of waiting for per-thread global loads.
  int nslots, zstage;
  int stage_off, xch_off;   // byte offsets of the ring and of the epilogue exchange buffer
  int xch_alias;            // the exchange buffer lives inside the B2 image (tight fits): one more worker barrier per plane
  int ablate;               // DENO_AN_ABLATE (measurement only, results WRONG): bit 0 no stage-1 MMAs, 1 no stage-2 MMAs, 2 no A-image stores,
                            // 3 no B2 stores, 4 no spectrum stores, 5 no gz stores, 6 no staged-plane loads in the transform
  int sync_off;             // byte offset of the barrier words (6 x 8 bytes) inside the dynamic allocation
};

struct AnalysisTcSmem {
  int e1, a2, a, b2, xch, stage, total;   // byte offsets
  int e1_bytes, a2_one, a_one, a_lbo, b2_bytes, b2_lbo;
};

constexpr int kAnWorkers = 512;               // 16 worker warps: every SIMT phase of the pipeline is latency-bound per warp
constexpr int kAnThreads = kAnWorkers + 32;   // + the MMA-issuing warp
constexpr int kAnZRegs = 5;                   // float4 of act'(z) per worker loaded ahead of the staged plane

__host__ __device__ inline AnalysisTcSmem analysis_tc_smem(int H, int W, int KH, int N1, int Wp, int Hp, int KRp, int nsub = 1) {
  AnalysisTcSmem m;
  m.e1_bytes = 2 * N1 * Wp * 4;
  m.a2_one = KRp * Hp * 4;
  m.a_lbo = 16 * Hp + 16;                   // padded K-direction strides: conflict-free stores when lanes walk along K
  m.a_one = (Wp / 4) * m.a_lbo;
  m.b2_lbo = 16 * 2 * N1 + 16;
  m.b2_bytes = (Hp / 4) * m.b2_lbo;
  int off = 0;
  m.e1 = off; off += m.e1_bytes;
  m.a2 = off; off += nsub * 2 * m.a2_one;   // [slab][hi | lo]
  m.a = off; off += 2 * m.a_one;
  m.b2 = off; off += m.b2_bytes;
  m.xch = off; off += round_up(KH * (N1 + 1) * 4, 16);
  m.stage = off; off += round_up(H * W, 4) * 4;
  // The M = 128 descriptors always read 16 row groups (2 KB) per core-matrix column, also for operands with
  // fewer rows: keep that over-read inside the allocation for tiny planes.
  m.total = off + 4096;
  return m;
}

__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kAnBarA = 2;         // A images stored (workers arrive, MMA warp syncs)
constexpr int kAnBarEpi0 = 8;      // 8..11: sin rows exchanged, one barrier per worker warp group
constexpr int kAnBarWorkers = 12;  // workers only
constexpr int kAnBarB2 = 13;       // B2 image stored (workers arrive, MMA warp syncs)

// Timing-only ablation bits (AnalysisTcParams::ablate) are compiled in only with -DDENO_AN_ABLATE_BUILD
// (tools/an_ablate.py builds that variant): in the product kernel the tests cost registers and issue slots.
#ifdef DENO_AN_ABLATE_BUILD
#define DENO_AN_ABL(bit) (p.ablate & (bit))
#else
#define DENO_AN_ABL(bit) 0
#endif
// EVEN: the plane width is even, i.e. every row of the staged slab starts 8-byte aligned (64-bit accesses, gz fused into
// the transform loop).  A template argument: as a run-time branch inside the transform loop it cost the odd-width Pipe
// planes 2 % of the layer (two code paths live in the hot loop of a latency-bound kernel).
// MODE: how slabs reach shared memory - 0: per-thread cp.async (sizes / offsets not multiples of 16 bytes), 1: one bulk copy,
// one slot, 2: two-slot ring, 3: one slot with act'(z) staged next to gy, 4: two-slot ring with act'(z) (3, 4: FUSE_GZ only).
// Template arguments for the same reason as EVEN.
// N1C: compile-time stage-1 N (0: run-time p.N1).  The launcher always passes 0: fixing N1 = 32 (every demo configuration)
// was measured SLOWER (3-D layer 1927 -> 1972 us, profiles/round2_ab_analysis_synthesis_specialisations.txt).
template <bool FUSE_GZ, bool SLABS, bool EVEN, int MODE, int N1C>
__global__ void __launch_bounds__(kAnThreads, 1) plane_analysis_tc_kernel(AnalysisTcParams p) {
  // no-swizzle descriptors need 16-byte alignment only; a 1024-byte aligned array would cost the tight fits 1 KB
  extern __shared__ __align__(128) unsigned char an_smem[];
  unsigned char* const smem = an_smem;
  // The barriers and the TMEM slot live at the end of the DYNAMIC allocation: static shared memory of a kernel in this
  // translation unit is padded to the 1024-byte alignment of the other kernels' arrays, which the tight fits cannot spare.
#ifdef DENO_AN_STATICBAR
  __shared__ __align__(8) uint64_t static_sync_words[6];
  uint64_t* const sync_words = static_sync_words;
#else
  uint64_t* const sync_words = reinterpret_cast<uint64_t*>(an_smem + p.sync_off);
#endif
  uint64_t* const mbar = sync_words;                   // [2]
  uint64_t* const bar_stage = sync_words + 2;          // [2]
  uint64_t& bar_const = sync_words[4];
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(sync_words + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A plane of up to 128 rows is one item.  Taller planes (the 229- and 137-row Airfoil / Pipe grids) are cut into
  // nsub row slabs of HS rows: every slab goes through the staging buffer, the transform and stage 1 like a plane of
  // its own, and stage 2 - the contraction over the rows - accumulates the slabs in TMEM with the slab's own row
  // twiddle image; the epilogue runs after the last slab.  Below, H is the slab height (HS) and `hrows` the rows of the
  // current slab (the last one may be shorter).
  const int Hfull = p.g.H, W = p.g.W, KH = p.g.KH, MW = p.g.MW;
  const int nsub = SLABS ? p.nsub : 1, H = p.HS;      // SLABS = false: the one-slab code is the compile-time special case
  const int N1 = N1C > 0 ? N1C : p.N1, Wp = p.Wp, Hp = p.Hp, KS0 = p.KS0, KRp = p.KRp;
  const AnalysisTcSmem so = analysis_tc_smem(H, W, KH, N1, Wp, Hp, KRp, nsub);
  const int HW = H * W;                                  // elements of a full slab
  const int hlast = Hfull - (nsub - 1) * H;              // rows of the last slab
  float* const stage_ring = reinterpret_cast<float*>(smem + p.stage_off);
  const int stage_floats = round_up((nsub == 1 ? Hfull : H) * W, 4);   // one staged slab (u, or act'(z)): the rows that exist
  constexpr bool bulk = MODE > 0;
  constexpr bool two_slots = MODE == 2 || MODE == 4;
  constexpr bool zstage = FUSE_GZ && MODE >= 3;
  const int slot_floats = stage_floats * (zstage ? 2 : 1);
  auto slot_of = [&](int j) { return two_slots ? (j & 1) : 0; };

  const uint32_t tmem_cols = 6 * N1 <= 64 ? 64u : (6 * N1 <= 128 ? 128u : (6 * N1 <= 256 ? 256u : 512u));
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 32) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_const, 1);
    mbar_fence_init();
    // constant operand images (final layout in global memory): two bulk async copies
    mbar_expect_tx(&bar_const, (uint32_t)(so.e1_bytes + nsub * 2 * so.a2_one));
    bulk_g2s(smem + so.e1, p.e1, (uint32_t)so.e1_bytes, &bar_const);
    bulk_g2s(smem + so.a2, p.a2, (uint32_t)(nsub * 2 * so.a2_one), &bar_const);
  }
  fence_before_sync();
  __syncthreads();                                     // barriers initialised, TMEM allocated
  fence_after_sync();
  DENO_PDL_SYNC();                                     // the planes (and act'(z)) come from the predecessor
  mbar_wait(&bar_const, 0);                            // constant images stored
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + (uint32_t)(3 * N1);   // [hi*hi | hi*lo | lo*hi] tiles per stage
  const uint32_t idesc_n2 = make_idesc_tf32(128, 2 * N1), idesc_n1 = make_idesc_tf32(128, N1);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t lbo_a = (uint32_t)so.a_lbo, lbo_e1 = 32 * N1, lbo_a2 = 16 * KRp, lbo_b2 = (uint32_t)so.b2_lbo;
  const int nyq = Wp / 4;                 // column quads of a row
  const int nchunk = (nyq + 7) / 8;       // a warp covers 8 quads (32 columns) x 4 rows per step
  // (handing the A images to the MMA warp in column groups, so that the stage-1 K loop chases the transform, was measured
  // slower at every shape - one more fence + barrier per group: profiles/round2_ab_analysis_chase.txt)
  const int nplanes_mine = (p.nplanes - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // planes of this CTA
  const int nmine = nplanes_mine * nsub;                 // items (slabs) of this CTA: item j = slab j % nsub of its plane j / nsub
  // Slabs whose byte size and offsets are multiples of 16 stream in with ONE bulk async copy (cp.async.bulk,
  // completion on an mbarrier) issued by the MMA warp; others fall back to per-thread 4-byte cp.async by the workers.
  // (MODE > 0 requires ((Hfull * W) % 4) == 0 && (HW % 4) == 0 && ((hlast * W) % 4) == 0: checked by the launcher)
  // (the one-slab case takes no integer division here: these run on the MMA warp's issuing lane, the critical path -
  // ten divisions per item cost 18 % at the 3-D turbulence shape with its 151 small planes per CTA)
  constexpr bool one_slab = !SLABS;
  auto item_sub = [&](int j) { return one_slab ? 0 : j % nsub; };
  auto item_plane = [&](int j) { return (int)blockIdx.x + (one_slab ? j : j / nsub) * (int)gridDim.x; };
  auto item_rows = [&](int j) { return (one_slab || item_sub(j) == nsub - 1) ? hlast : H; };
  auto item_base = [&](int j) {                          // element offset of slab j of this CTA
    const int q = item_plane(j);
    return plane_base(p.g, q / p.g.C, q % p.g.C) + (long long)item_sub(j) * HW;
  };

  if (warp == kAnWorkers / 32) {
    // ------------------------------------------------------------------ MMA warp
    const bool leader = elect_one();
    auto l2_prefetch = [&](int j) {                    // pull slab j (and its act'(z)) into L2 ahead of its bulk copy
      if (!bulk || j >= nmine || !leader) return;
      const long long base = item_base(j);
      const uint32_t bytes = (uint32_t)(item_rows(j) * W) * 4u;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.u + base), "r"(bytes) : "memory");
      if (FUSE_GZ) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.z + base), "r"(bytes) : "memory");
    };
    auto prefetch_plane = [&](int j) {                 // bulk copy of this CTA's j-th slab (and its act'(z)) into ring slot j % nslots
      if (!bulk || j >= nmine || !leader) return;
      const long long base = item_base(j);
      const uint32_t bytes = (uint32_t)(item_rows(j) * W) * 4u;
      const int slot = slot_of(j);
      const uint32_t bar = smem_u32(&bar_stage[slot]);
      float* dst = stage_ring + slot * slot_floats;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(zstage ? 2u * bytes : bytes) : "memory");
      // (issuing a slab as several smaller bulk copies was measured slower: profiles/round2_ab_analysis_piece_copies.txt)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dst)), "l"(p.u + base), "r"(bytes), "r"(bar) : "memory");
      if (zstage)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + stage_floats)), "l"(p.z + base), "r"(bytes), "r"(bar) : "memory");
    };
    long long idle = 0;
    auto stage1 = [&](int j) {                         // stage-1 MMAs of plane j once its A images are stored
      uint64_t ah = make_desc(sbase + so.a, lbo_a, 128), al = make_desc(sbase + so.a + so.a_one, lbo_a, 128);
      uint64_t be = make_desc(sbase + so.e1, lbo_e1, 128);
      const uint64_t inc_a = (uint64_t)((2 * lbo_a) >> 4), inc_b = (uint64_t)((2 * lbo_e1) >> 4);
      const long long t0 = DENO_CLK();
      nbar_sync(kAnBarA, kAnThreads);
      fence_after_sync();
      idle += DENO_CLK() - t0;
      // With a two-slot ring the MMAs go first: the workers wait for them next, while the copy below is for the item
      // after next and its address arithmetic (integer divisions in item_base) is ~0.5 k cycles of this single issuing
      // lane.  With one slot the copy is what the next transform waits for, so it keeps its place in front.
      if (!two_slots) { prefetch_plane(j + 1); l2_prefetch(j + 2); }
      for (int s2 = 0; s2 < Wp / 8; ++s2) {
        const uint32_t acc = s2 > 0 ? 1u : 0u;
        if (leader && !DENO_AN_ABL(1)) {
          mma_tf32(tmem_d1, ah, be, idesc_n2, acc);              // hi * [hi | lo]
          mma_tf32(tmem_d1 + 2 * N1, al, be, idesc_n1, acc);     // lo * hi
        }
        ah += inc_a; al += inc_a; be += inc_b;
      }
      if (leader) mma_commit(&mbar[0]);
      if (two_slots) { prefetch_plane(j + 2); l2_prefetch(j + 3); }   // every worker is past item j's ring slot
      __syncwarp();
    };
    prefetch_plane(0);
    if (two_slots) prefetch_plane(1);
    l2_prefetch(p.nslots);
    stage1(0);
    for (int j = 0; j < nmine; ++j) {
      const long long t0 = DENO_CLK();
      nbar_sync(kAnBarB2, kAnThreads);                 // B2(j) stored; D1 and D2 of the previous plane drained
      fence_after_sync();
      idle += DENO_CLK() - t0;
      const int sub = item_sub(j);                     // slab: its row-twiddle image; slabs after the first accumulate
      const uint32_t a2base = sbase + so.a2 + sub * 2 * so.a2_one;
      uint64_t a2h = make_desc(a2base, lbo_a2, 128), a2l = make_desc(a2base + so.a2_one, lbo_a2, 128);
      uint64_t b2 = make_desc(sbase + so.b2, lbo_b2, 128);
      const uint64_t inc_a2 = (uint64_t)((2 * lbo_a2) >> 4), inc_b2 = (uint64_t)((2 * lbo_b2) >> 4);
      for (int s2 = 0; s2 < Hp / 8; ++s2) {
        const uint32_t acc = (s2 > 0 || sub > 0) ? 1u : 0u;
        if (leader && !DENO_AN_ABL(2)) {
          mma_tf32(tmem_d2, a2h, b2, idesc_n2, acc);
          mma_tf32(tmem_d2 + 2 * N1, a2l, b2, idesc_n1, acc);
        }
        a2h += inc_a2; a2l += inc_a2; b2 += inc_b2;
      }
      if (leader) mma_commit(&mbar[1]);
      __syncwarp();
      if (j + 1 < nmine) stage1(j + 1);
    }
    if (leader) DENO_DBG_ADD(34, idle);                // slot 34: MMA warp waiting for the workers
  } else {
    // ------------------------------------------------------------------ workers
    const int wq = warp & 3, wg = warp >> 2;
    const int xrow = wq * 32 + lane;        // TMEM lane owned by this thread (row x of D1 / row r of D2)
    uint32_t phase = 0;
    float* const stage = stage_ring;                   // non-bulk fallback: one slot
    auto issue_plane_cp_async = [&](int j) {           // non-bulk fallback
      const long long base = item_base(j);
      const int n = item_rows(j) * W;
      for (int i = tid; i < n; i += kAnWorkers) cp_async4(stage + i, p.u + base + i);
      cp_async_commit();
    };
    auto transform = [&](int j) {                      // staged slab j -> (gz) -> A images (hi / lo)
      const int hrows = item_rows(j);
      const int nel = hrows * W;
      const long long tq0 = DENO_CLK();
      float* const stage = stage_ring + slot_of(j) * slot_floats;
      float4 zr[kAnZRegs];                             // act'(z) of this thread's first elements: in flight during the wait
      if (FUSE_GZ && bulk && !zstage) {
        const float4* z4 = reinterpret_cast<const float4*>(p.z + item_base(j));
#pragma unroll
        for (int r = 0; r < kAnZRegs; ++r) {
          const int i = tid + r * kAnWorkers;
          zr[r] = i < nel / 4 ? __ldg(z4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (bulk) {
        mbar_wait(&bar_stage[slot_of(j)], (uint32_t)((two_slots ? (j >> 1) : j) & 1));
      } else {
        cp_async_wait_all();
        nbar_sync(kAnBarWorkers, kAnWorkers);
      }
      const long long tq1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(32, tq1 - tq0);      // slot 32: wait for the staged plane
      // gz = gy * act'(z).  With act'(z) staged next to gy the product is formed INSIDE the transform loop below: the
      // thread that splits a row quad also multiplies it and stores it to gz (a warp's store covers 4 row segments of
      // 128 bytes), so there is no separate pass over the plane, no write-back into the staging buffer and no barrier
      // between the two (that pass + barrier was 1.5 - 2.3 k cycles per plane, profiles/round2_s3_phase_probe.txt).
      constexpr bool fuse_in_loop = EVEN && zstage;    // odd widths: rows are 4-byte aligned only, the float4 pass below is faster
      if (FUSE_GZ && !fuse_in_loop) {                  // separate pass over the slab, gz kept in staging
        const long long base = item_base(j);
        if (zstage) {                                  // both factors staged: shared memory -> shared memory + global
          float4* s4 = reinterpret_cast<float4*>(stage);
          const float4* zs4 = reinterpret_cast<const float4*>(stage + stage_floats);
          float4* g4 = reinterpret_cast<float4*>(p.gz_out + base);
          for (int i = tid; i < nel / 4; i += kAnWorkers) {
            const float4 a = s4[i], b = zs4[i];
            const float4 v = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
            s4[i] = v;
            if (!DENO_AN_ABL(32)) g4[i] = v;
          }
        } else if (bulk) {
          float4* s4 = reinterpret_cast<float4*>(stage);
          float4* g4 = reinterpret_cast<float4*>(p.gz_out + base);
#pragma unroll
          for (int r = 0; r < kAnZRegs; ++r) {
            const int i = tid + r * kAnWorkers;
            if (i < nel / 4) {
              const float4 a = s4[i], b = zr[r];
              const float4 v = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
              s4[i] = v;
              g4[i] = v;
            }
          }
          const float4* z4 = reinterpret_cast<const float4*>(p.z + base);
          for (int i = tid + kAnZRegs * kAnWorkers; i < nel / 4; i += kAnWorkers) {   // planes beyond the register window
            const float4 a = s4[i], b = __ldg(z4 + i);
            const float4 v = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
            s4[i] = v;
            g4[i] = v;
          }
        } else {
          for (int i0 = tid; i0 < nel; i0 += 8 * kAnWorkers) {      // eight act'(z) loads in flight per thread
            float zz[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) zz[u] = ldg_ordered(p.z + base + min(i0 + u * kAnWorkers, nel - 1));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = i0 + u * kAnWorkers;
              if (i < nel) {
                const float v = stage[i] * zz[u];
                stage[i] = v;
                p.gz_out[base + i] = v;
              }
            }
          }
        }
        nbar_sync(kAnBarWorkers, kAnWorkers);
      }
      const long long tq2 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(40, tq2 - tq1);      // slot 40: gz pass
      {
        const int r0 = warp * 4 + (lane >> 3), r1 = r0 + (kAnWorkers / 32) * 4;     // Hp <= 128: at most two rows per thread
        const float* s0 = stage + r0 * W;
        const float* s1 = stage + r1 * W;
        constexpr bool fuse = fuse_in_loop;            // multiply by the staged act'(z) and store gz here
        const float* zs = stage + stage_floats;
        float* gout = (FUSE_GZ && EVEN) ? p.gz_out + item_base(j) : nullptr;
        constexpr bool even = EVEN;                    // rows start 8-byte aligned: 64-bit shared / global accesses
        // one row quad (4 consecutive columns of row r): load, (multiply + store gz), zero outside the slab
        auto load_quad = [&](const float* srow, int r, int y0, float (&v)[4]) {
          const bool rok = r < hrows && !DENO_AN_ABL(64);
          {
            float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
            if (rok && y0 < W) a = *reinterpret_cast<const float2*>(srow + y0);
            if (rok && y0 + 2 < W) b = *reinterpret_cast<const float2*>(srow + y0 + 2);
            if (fuse) {
              const float* zrow = zs + (srow - stage);
              float* grow = gout + (srow - stage);
              if (rok && y0 < W) {
                const float2 za = *reinterpret_cast<const float2*>(zrow + y0);
                a.x *= za.x; a.y *= za.y;
                if (!DENO_AN_ABL(32)) *reinterpret_cast<float2*>(grow + y0) = a;
              }
              if (rok && y0 + 2 < W) {
                const float2 zb = *reinterpret_cast<const float2*>(zrow + y0 + 2);
                b.x *= zb.x; b.y *= zb.y;
                if (!DENO_AN_ABL(32)) *reinterpret_cast<float2*>(grow + y0 + 2) = b;
              }
            }
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
          }
        };
        for (int c = 0; c < nchunk; ++c) {
          const int yq = c * 8 + (lane & 7);
          if (yq < nyq) {
            float v0[4], v1[4];
            if (even) {
              load_quad(s0, r0, yq * 4, v0);
              load_quad(s1, r1, yq * 4, v1);
            } else {                                   // odd widths (137-wide Pipe planes): plain predicated 4-byte loads
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int y = yq * 4 + jj;
                v0[jj] = (r0 < hrows && y < W) ? s0[y] : 0.f;
                v1[jj] = (r1 < hrows && y < W) ? s1[y] : 0.f;
              }
            }
            float4 hi, lo;
            if (r0 < Hp && !DENO_AN_ABL(4)) {
              split_tf32(v0[0], hi.x, lo.x);
              split_tf32(v0[1], hi.y, lo.y);
              split_tf32(v0[2], hi.z, lo.z);
              split_tf32(v0[3], hi.w, lo.w);
              const int off = (r0 & 7) * 16 + (r0 >> 3) * 128 + yq * so.a_lbo;
              *reinterpret_cast<float4*>(smem + so.a + off) = hi;
              *reinterpret_cast<float4*>(smem + so.a + so.a_one + off) = lo;
            }
            if (r1 < Hp && !DENO_AN_ABL(4)) {
              split_tf32(v1[0], hi.x, lo.x);
              split_tf32(v1[1], hi.y, lo.y);
              split_tf32(v1[2], hi.z, lo.z);
              split_tf32(v1[3], hi.w, lo.w);
              const int off = (r1 & 7) * 16 + (r1 >> 3) * 128 + yq * so.a_lbo;
              *reinterpret_cast<float4*>(smem + so.a + off) = hi;
              *reinterpret_cast<float4*>(smem + so.a + so.a_one + off) = lo;
            }
          }
        }
      }
      fence_proxy_async();
      nbar_arrive(kAnBarA, kAnThreads);
      if (!bulk && j + 1 < nmine) {
        nbar_sync(kAnBarWorkers, kAnWorkers);          // every worker is past the staging buffer
        issue_plane_cp_async(j + 1);
      }
      if (warp == 1) { DENO_DBG_ADD(33, DENO_CLK() - tq1); DENO_DBG_ADD(41, DENO_CLK() - tq2); }   // slots 33/41: (gz +) transform, chunk loop
    };

    if (!bulk) issue_plane_cp_async(0);
    transform(0);
    for (int j = 0; j < nmine; ++j) {
      const int q = item_plane(j);
      const int hrows = item_rows(j);
      const bool last_slab = one_slab || item_sub(j) == nsub - 1;
      const long long tq3 = DENO_CLK();
      mbar_wait(&mbar[0], phase);
      fence_after_sync();
      const long long tq4 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(35, tq4 - tq3);      // slot 35: wait for the stage-1 MMAs
      // D1 -> B2 image [n][x]: hi rows n, lo rows N1 + n
      {
        const int ncol_q = N1 / 4;                 // columns per warp group (N1 is a multiple of 32)
        unsigned char* b2hi = smem + so.b2;
        unsigned char* b2lo = b2hi + (N1 >> 3) * 128;
        for (int c0 = 0; c0 < ncol_q; c0 += 8) {
          float v[8], v1[8], v2[8];
          const uint32_t ta = tmem_d1 + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wg * ncol_q + c0);
          tmem_ld8(ta, v);
          tmem_ld8(ta + N1, v1);
          tmem_ld8(ta + 2 * N1, v2);
          if (xrow < Hp) {
            const int kc = (xrow & 3) * 4 + (xrow >> 2) * so.b2_lbo;     // k' = x
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int n = wg * ncol_q + c0 + jj;
              const float t = xrow < hrows ? v[jj] + (v1[jj] + v2[jj]) : 0.f;
              float th, tl;
              split_tf32(t, th, tl);
              const int r0 = (n & 7) * 16 + (n >> 3) * 128;
              if (!DENO_AN_ABL(8)) {
                *reinterpret_cast<float*>(b2hi + r0 + kc) = th;
                *reinterpret_cast<float*>(b2lo + r0 + kc) = tl;
              }
            }
          }
        }
      }
      fence_proxy_async();
      fence_before_sync();
      nbar_arrive(kAnBarB2, kAnThreads);
      const long long tq5 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(36, tq5 - tq4);      // slot 36: D1 -> B2 conversion
      if (j + 1 < nmine) transform(j + 1);
      const long long tq6 = DENO_CLK();
      mbar_wait(&mbar[1], phase);
      fence_after_sync();
      const long long tq7 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(37, tq7 - tq6);      // slot 37: wait for the stage-2 MMAs
      // epilogue (after the plane's last slab): rows k (cos) and KS0 + k (sin) of D2 are combined through shared
      // memory; warp group wg takes the column passes c0 = 16 wg, 16 wg + 64, ...
      if (last_slab) {
        float* xch = reinterpret_cast<float*>(smem + p.xch_off);  // [KH][N1 + 1] sin-row values (odd pitch: conflict-free)
        const bool is_cos = xrow < KH, is_sin = xrow >= KS0 && xrow < KS0 + KH;
        const bool warp_used = (wq * 32 < KH) || (wq * 32 + 32 > KS0 && wq * 32 < KS0 + KH);   // any cos / sin row in this quadrant
        for (int c0 = wg * 16; c0 < N1; c0 += 64) {
          float dsum[16];
          if (warp_used) tmem_ld16x3_sum(tmem_d2 + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, (uint32_t)N1, dsum);
          if (is_sin) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) xch[(xrow - KS0) * (N1 + 1) + c0 + jj] = dsum[jj];
          }
          nbar_sync(kAnBarEpi0 + wg, 128);
          if (is_cos) {
            float2* dst = p.out + (long long)q * KH * MW + (long long)xrow * MW;
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) {
              const int l = (c0 + jj) >> 1;
              if (l < MW) {
                const float sr = xch[xrow * (N1 + 1) + c0 + jj], si = xch[xrow * (N1 + 1) + c0 + jj + 1];
                const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
                if (!DENO_AN_ABL(16)) dst[l] = make_float2((dsum[jj] + si) * f, (dsum[jj + 1] - sr) * f);
              }
            }
          }
        }
      }
      fence_before_sync();                             // D2 reads ordered before this thread's next hand-off
      if (p.xch_alias) nbar_sync(kAnBarWorkers, kAnWorkers);   // exchange buffer inside B2: no worker may store B2(j+1) before all have read it
      if (warp == 1) { DENO_DBG_ADD(38, DENO_CLK() - tq7); DENO_DBG_ADD(39, 1); }   // slots 38/39: epilogue, planes
      phase ^= 1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// ----------------------------------------------------------------------------------------------
// bypass_wgrad_tc: gradient of the 1x1 bypass conv weight as one long-K GEMM on the tensor cores.
//   gW[o, i] = sum_{b, pt} gz[b, o, pt] * x[b, i, pt]        (K = points; both operands K-major as stored)
// A tcgen05.mma costs the same ~74 cycles for every N <= 128 (tools/tc_mma_bench.cu), and a single
// 32 x 32 product would use 6 % of a 128 x 128 instruction, so NS batch entries are stacked along BOTH
// M and N:  D[(j,o), (j',i)] = sum_pt gz[b_j, o, pt] * x[b_j', i, pt];  only the diagonal blocks j = j'
// are wanted, but the instruction count drops NS-fold.  Persistent CTAs accumulate over their tiles of
// KT = 64 points in TMEM (three split-pass tiles) and write NS partials (CO x CI) each;
// bypass_wgrad_reduce_kernel sums the partials in a fixed order.  The bias gradient comes from the DC
// mode of the packed spectrum G^ (bias_from_dc_kernel), which is exactly sum_pt gz.
// ----------------------------------------------------------------------------------------------
struct WgradTcParams {
  PlaneGeom gx;          // geometry of x (C = CI)
  int CO;
  long long sbo_g;       // bo stride of gz
  const float* gz;
  const float* x;
  float* partial;        // [gridDim.x * NS][CO*CI]
  int NS;                // stacked batch entries per tile
  int COs, CIs;          // per-slot row counts (CO, CI rounded up to 8)
  int tiles_per_plane;   // ceil(H*W / 64)
  int ngroups;           // ceil(NB' / NS)
  int ntiles;            // ngroups * tiles_per_plane
  int nbatch;            // NB'
  int vec4;              // H*W % 4 == 0: 16-byte global loads are aligned
  int gz_tps;            // > 0: gz is stored in blocks of 128 points, [batch][gz_tps blocks][CO][128] (fno_proj_bwd_tc)
};

constexpr int kWgradKT = 32;        // points per tile (two shared-memory stages of operand images)
constexpr int kWgradLoaders = 512;  // loader threads (16 warps): bytes in flight per SM are what bounds this kernel
constexpr int kWgradThreads = kWgradLoaders + 32;  // + 1 MMA warp

struct WgradTcSmem {
  int stage[2];          // byte offset of each stage: [g_hi | g_lo | x_hi | x_lo]
  int g_lo, x_hi, x_lo;  // offsets inside a stage (g_hi = 0)
  int total;
  int g_lbo, x_lbo;
};

__host__ __device__ inline WgradTcSmem wgrad_tc_smem(int NS, int COs, int CIs) {
  WgradTcSmem m;
  const int mrows = 128;                 // the M = 128 descriptor always reads 16 row groups
  const int nrows = NS * CIs;
  m.g_lbo = 16 * mrows + 16;             // padded K-direction strides: conflict-free 16-byte stores
  m.x_lbo = 16 * nrows + 16;
  const int g_one = (kWgradKT / 4) * m.g_lbo, x_one = (kWgradKT / 4) * m.x_lbo;
  m.g_lo = g_one;
  m.x_hi = 2 * g_one;
  m.x_lo = 2 * g_one + x_one;
  const int stage_bytes = 2 * g_one + 2 * x_one;
  m.stage[0] = 0;
  m.stage[1] = stage_bytes;
  m.total = 2 * stage_bytes + 256;
  (void)COs;
  return m;
}

// 8 loader warps stream (row, point-quad) items global -> registers (two tiles in flight) -> tf32 split ->
// operand images (two stages); warp 8 issues the MMAs.  Stages are handed over with mbarriers
// (full: 256 loader arrivals; empty: tcgen05.commit), so loads, splitting and tensor-core work overlap.
template <int NIT>   // items per loader thread and tile: NS * (CO + CI) * (kWgradKT / 4) / kWgradLoaders rounded up
__global__ void __launch_bounds__(kWgradThreads, 1) bypass_wgrad_tc_kernel(WgradTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CI = p.gx.C, CO = p.CO, NS = p.NS, COs = p.COs, CIs = p.CIs;
  const WgradTcSmem so = wgrad_tc_smem(NS, COs, CIs);
  const int HW = p.gx.H * p.gx.W;
  const int rows_per_slot = CO + CI;
  const int nrows = NS * rows_per_slot;      // rows streamed per tile
  const int N = NS * CIs;
  constexpr int QPR = kWgradKT / 4;          // point quads per row
  uint32_t ncols = 32;
  while ((int)ncols < 3 * N) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    for (int s2 = 0; s2 < 2; ++s2) {
      mbar_init(&bar_full[s2], kWgradLoaders);
      mbar_init(&bar_empty[s2], 1);
    }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < so.total / 16; i += kWgradThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  DENO_PDL_SYNC();                                     // gz and x come from earlier kernels
  const uint32_t tmem = tmem_slot;
  // tiles of this CTA: t = blockIdx.x + n * gridDim.x
  const int ntile_cta = (p.ntiles > (int)blockIdx.x) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < kWgradLoaders / 32) {
    // ===== loader warps =====
    // The (row, point-quad) -> (slot, channel, image offset) map does not depend on the tile: decode it once.
    // meta: bit 31 valid, bit 30 "gz row", bits 28..29 slot j, bits 0..27 byte offset inside the hi image.
    uint32_t meta[NIT];
    int roff[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int item = it * kWgradLoaders + tid;
      const int row = item / QPR, pq = item - row * QPR;
      meta[it] = 0u;
      roff[it] = 0;
      if (row < nrows) {
        const int j = row / rows_per_slot, r = row - j * rows_per_slot;
        const bool isg = r < CO;
        const int ir = isg ? j * COs + r : j * CIs + (r - CO);
        const int off = (ir & 7) * 16 + (ir >> 3) * 128 + pq * (isg ? so.g_lbo : so.x_lbo);
        meta[it] = 0x80000000u | (isg ? 0x40000000u : 0u) | ((uint32_t)j << 28) | (uint32_t)off;
        roff[it] = (isg && p.gz_tps > 0) ? r * 128 + pq * 4 : (int)((isg ? r : r - CO) * p.gx.sc) + pq * 4;
      }
    }
    float4 ra[NIT], rb[NIT];
    // Per tile and item only the batch entry's base offset changes: keep the address arithmetic to one or two
    // 64-bit multiply-adds per item (an earlier version spent ~1200 instructions per warp and tile on integer
    // divisions and selects and was issue-bound on them, profiles/r2n_ncu_C2_layer_kernels.csv).
    const bool flat_batch = p.gx.NBI == 1;
    const long long blk_stride = (long long)CO * 128;
    auto prefetch = [&](int n, float4 (&regs)[NIT]) {
      const int t = (int)blockIdx.x + n * (int)gridDim.x;
      const int grp = t / p.tiles_per_plane;
      const int p0 = (t - grp * p.tiles_per_plane) * kWgradKT;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t m = meta[it];
        const int bp = grp * NS + (int)((m >> 28) & 3u);
        if ((m & 0x80000000u) && bp < p.nbatch) {
          const bool isg = (m & 0x40000000u) != 0u;
          int bo = bp, bi = 0;
          if (!flat_batch) {
            bo = bp / p.gx.NBI;
            bi = bp - bo * p.gx.NBI;
          }
          long long base;
          if (isg && p.gz_tps > 0) base = ((long long)bp * p.gz_tps + (p0 >> 7)) * blk_stride + (p0 & 127);
          else base = (long long)bo * (isg ? p.sbo_g : p.gx.sbo) + (long long)bi * p.gx.sbi + p0;
          const float* src = (isg ? p.gz : p.x) + base + roff[it];
          const int pt = p0 + ((it * kWgradLoaders + tid) % QPR) * 4;
          if (p.vec4 && pt + 3 < HW) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (pt + 0 < HW) v.x = src[0];
            if (pt + 1 < HW) v.y = src[1];
            if (pt + 2 < HW) v.z = src[2];
            if (pt + 3 < HW) v.w = src[3];
          }
        }
        regs[it] = v;
      }
    };
    auto publish = [&](int n, float4 (&regs)[NIT]) {
      const int st = n & 1;
      const long long tp0 = DENO_CLK();
      mbar_wait(&bar_empty[st], (uint32_t)(((n >> 1) & 1) ^ 1));      // MMAs of tile n-2 are done with this stage
      const long long tp1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(10, tp1 - tp0);                      // slot 10: loaders waiting for a free stage
      unsigned char* base = smem + so.stage[st];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const uint32_t m = meta[it];
        if (m & 0x80000000u) {
          float4 hi, lo;
          split_tf32(regs[it].x, hi.x, lo.x);
          split_tf32(regs[it].y, hi.y, lo.y);
          split_tf32(regs[it].z, hi.z, lo.z);
          split_tf32(regs[it].w, hi.w, lo.w);
          const bool isg = (m & 0x40000000u) != 0u;
          unsigned char* dst = base + (isg ? 0 : so.x_hi) + (m & 0x0FFFFFFFu);
          *reinterpret_cast<float4*>(dst) = hi;
          *reinterpret_cast<float4*>(dst + (isg ? so.g_lo : so.x_lo - so.x_hi)) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[st]);
      if (warp == 1) { DENO_DBG_ADD(11, DENO_CLK() - tp1); DENO_DBG_ADD(15, 1); }   // slots 11/15: wait for the data + split + stores, tiles
    };
    if (ntile_cta > 0) prefetch(0, ra);
    if (ntile_cta > 1) prefetch(1, rb);
    for (int n = 0; n < ntile_cta; n += 2) {   // (a third register buffer issued ahead of publish() spilled and was slower)
      publish(n, ra);
      long long tf0 = DENO_CLK();
      if (n + 2 < ntile_cta) prefetch(n + 2, ra);
      if (warp == 1) DENO_DBG_ADD(12, DENO_CLK() - tf0);                // slot 12: address arithmetic + load issue
      if (n + 1 < ntile_cta) {
        publish(n + 1, rb);
        tf0 = DENO_CLK();
        if (n + 3 < ntile_cta) prefetch(n + 3, rb);
        if (warp == 1) DENO_DBG_ADD(12, DENO_CLK() - tf0);
      }
    }
  } else {
    // ===== MMA warp =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint32_t sbase = smem_u32(smem);
    const uint64_t inc_g = (uint64_t)((2 * so.g_lbo) >> 4), inc_x = (uint64_t)((2 * so.x_lbo) >> 4);
    for (int n = 0; n < ntile_cta; ++n) {
      const int st = n & 1;
      const long long tm0 = DENO_CLK();
      mbar_wait(&bar_full[st], (uint32_t)((n >> 1) & 1));
      fence_after_sync();
      const long long tm1 = DENO_CLK();
      const uint32_t b0 = sbase + so.stage[st];
      uint64_t gh = make_desc(b0, so.g_lbo, 128), gl = make_desc(b0 + so.g_lo, so.g_lbo, 128);
      uint64_t xh = make_desc(b0 + so.x_hi, so.x_lbo, 128), xl = make_desc(b0 + so.x_lo, so.x_lbo, 128);
#pragma unroll
      for (int s2 = 0; s2 < kWgradKT / 8; ++s2) {
        const uint32_t acc = (n == 0 && s2 == 0) ? 0u : 1u;
        if (leader) {
          mma_tf32(tmem, gh, xh, idesc, acc);              // hi * hi
          mma_tf32(tmem + N, gl, xh, idesc, acc);          // lo * hi
          mma_tf32(tmem + 2 * N, gh, xl, idesc, acc);      // hi * lo
        }
        gh += inc_g; gl += inc_g; xh += inc_x; xl += inc_x;
      }
      if (leader) mma_commit(&bar_empty[st]);
      __syncwarp();
      if (leader) { DENO_DBG_ADD(13, tm1 - tm0); DENO_DBG_ADD(14, DENO_CLK() - tm1); }   // slots 13/14: MMA warp waiting / issuing
    }
    if (leader) mma_commit(&bar_done);
    __syncwarp();
  }
  // every warp: wait for the last MMAs, then warps 0-3 turn the accumulators into partials
  if (ntile_cta > 0) mbar_wait(&bar_done, 0);
  fence_after_sync();
  // lane (j, o) of TMEM holds row o of slot j; its diagonal block sits at columns j*CIs .. j*CIs + CI
  if (warp < 4) {
    const int row = (warp & 3) * 32 + lane;
    const int j = row / COs, o = row - j * COs;
    const bool valid = j < NS && o < CO;
    float* out = p.partial + ((long long)blockIdx.x * NS + (j < NS ? j : 0)) * (CO * CI) + (long long)o * CI;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16], v1[16], v2[16];
      const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
      tmem_ld16(ta, v);
      tmem_ld16(ta + N, v1);
      tmem_ld16(ta + 2 * N, v2);
      if (valid) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int i = c0 + jj - j * CIs;
          if (i >= 0 && i < CI) out[i] = ntile_cta > 0 ? v[jj] + (v1[jj] + v2[jj]) : 0.f;
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

// glb[o] = S * sum_b Re G^[b, o, mode 0]  (the DC mode of the scaled spectrum is sum_pt gz / S)
struct BiasDcParams { const float2* ghat; float* glb; int B, CO, M; float s_total; };
__global__ void bias_from_dc_kernel(BiasDcParams p) {
  DENO_PDL_SYNC();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= p.CO) return;
  float acc = 0.f;
  for (int b0 = 0; b0 < p.B; b0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v[u] = ldg_ordered(reinterpret_cast<const float*>(p.ghat + ((long long)min(b0 + u, p.B - 1) * p.CO + o) * p.M));
      if (b0 + u >= p.B) v[u] = 0.f;
    }
    DENO_SCHED_FENCE();
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  p.glb[o] = acc * p.s_total;
}

}  // namespace tc
}  // namespace deno
