// spectral_tma.cuh - tensor-map TMA + SWIZZLE_128B tcgen05 kernels of the spectral path (sm_100a only).
//
// The kernels of spectral_tc.cuh stage every tf32 operand through registers: global -> registers -> hi / lo split
// -> K-major no-swizzle image.  Where the operand is ALREADY K-major in HBM (the contraction index is the
// contiguous one) that detour is unnecessary: a tensor-map TMA copy (cp.async.bulk.tensor) drops a
// [rows][32 floats] box straight into shared memory in the canonical K-major SWIZZLE_128B layout that a tcgen05
// shared-memory descriptor reads (8-row atoms of 1 KB, 16-byte chunk index XOR row index; SBO = 1024).
//
// 3xTF32 on top of that: the tensor core ignores the low 13 mantissa bits of an fp32 operand, so the landed tile,
// masked IN PLACE to its tf32 part, is the "hi" operand, and a SIMT pass writes lo = v - hi into a second tile with
// the same layout.  That pass is element-wise, hence layout-agnostic: it never needs to know the swizzle.
#pragma once
#include <cuda.h>

#include "spectral_tc.cuh"

namespace deno {
namespace tc {

// cp.async.bulk.tensor 2-D tile load; coordinates are element indices {inner, outer}; out-of-bounds elements are
// zero-filled and still counted in the transaction bytes.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// K-major SWIZZLE_128B shared-memory descriptor of a [rows][32 fp32] tile (1024-byte aligned): 8-row atoms 1 KB apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused for a K-major swizzled tile one atom wide
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                 // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

// ----------------------------------------------------------------------------------------------
// bypass_wgrad_tma: gW[o, i] = sum_{b, pt} gz[b, o, pt] * x[b, i, pt]  (same contract as bypass_wgrad_tc_kernel).
// Both operands are K-major as stored (K = points).  Per tile of 32 points and NS stacked batch entries:
//   TMA     : NS boxes [COs rows][32] of gz -> A tile (128 rows), NS boxes [CIs rows][32] of x -> B tile (N rows)
//   builders: mask the landed tiles to tf32 in place (hi), write lo = v - hi tiles
//   MMA warp: per K = 8 step  D0 += A_hi B_hi,  D1 += A_lo B_hi,  D2 += A_hi B_lo   (three TMEM accumulators)
// Persistent CTAs (one per SM) accumulate over their tiles; at the end the NS diagonal blocks of D0 + D1 + D2 are
// summed inside the CTA and ONE partial (CO x CI) per CTA is written; bypass_wgrad_reduce_kernel adds the <= 148
// partials in a fixed order (bitwise reproducible).
// ----------------------------------------------------------------------------------------------
struct WgradTmaParams {
  float* partial;        // [gridDim.x][CO * CI]
  int CO, CI, COs, CIs, NS;
  int nchunks;           // ceil(S / 32)
  int ngroups;           // ceil(NB / NS)
  int ntiles;            // ngroups * nchunks
  int a_rows_per_batch;  // row index of batch entry b in the gz tensor map = b * a_rows_per_batch (CO)
  int b_rows_per_batch;  // ... in the x tensor map (CI)
  int a_tps;             // > 0: gz is stored in blocks of 128 points [batch][a_tps][CO][128] (fno_proj_bwd_tc)
  int nstages;
  int nbatch;            // batch entries; slots of the last group beyond it are never loaded (builders write zeros)
};

constexpr int kWtBuilders = 256;                       // 8 builder / epilogue warps
constexpr int kWtThreads = kWtBuilders + 64;           // + TMA warp + MMA warp
constexpr int kWtBarEpi = 3;                           // named barrier of the builder warps (epilogue exchange)

__host__ __device__ inline int wgrad_tma_stage_bytes(int N) { return 2 * 128 * 128 + 2 * N * 128; }   // A hi, A lo, B hi, B lo

__global__ void __launch_bounds__(kWtThreads, 1)
bypass_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, WgradTmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_tma[4], bar_lo[4], bar_empty[4], bar_done;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CO = p.CO, CI = p.CI, COs = p.COs, CIs = p.CIs, NS = p.NS;
  const int N = NS * CIs;
  const int NST = p.nstages;
  const int stage_bytes = wgrad_tma_stage_bytes(N);
  const int a_bytes = 128 * 128, b_bytes = N * 128;
  uint32_t ncols = 32;
  while ((int)ncols < 3 * N) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bar_tma[s], 1);
      mbar_init(&bar_lo[s], kWtBuilders / 32);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  // rows of the A tile that no box covers (NS * COs < 128) are still read by the M = 128 instruction: keep them finite
  for (int i = tid; i < NST * stage_bytes / 16; i += kWtThreads) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  DENO_PDL_SYNC();                                     // gz and x come from earlier kernels
  const uint32_t tmem = tmem_slot;
  const int ntile_cta = (p.ntiles > (int)blockIdx.x) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int n = 0; n < ntile_cta; ++n) {
        const int st = n % NST, use = n / NST;
        mbar_wait(&bar_empty[st], (uint32_t)((use & 1) ^ 1));     // MMAs of the previous use of this stage have retired
        const int t = (int)blockIdx.x + n * (int)gridDim.x;
        const int grp = t / p.nchunks, chunk = t - grp * p.nchunks;
        const int p0 = chunk * 32;
        unsigned char* sa = smem + st * stage_bytes;
        unsigned char* sb = sa + 2 * a_bytes;
        // Slots beyond the batch are NOT loaded: a box that lies entirely outside the tensor was measured to corrupt
        // other slots of the tile (profiles/r2c_wgrad_diag.txt); the builders write zeros there instead.
        const int nvalid = min(NS, p.nbatch - grp * NS);
        mbar_expect_tx(&bar_tma[st], (uint32_t)(nvalid * (COs + CIs) * 128));
        for (int j = 0; j < nvalid; ++j) {
          const int b = grp * NS + j;
          if (p.a_tps > 0) tma_load_2d(sa + j * COs * 128, &tm_a, p0 & 127, (b * p.a_tps + (p0 >> 7)) * p.a_rows_per_batch, &bar_tma[st]);
          else tma_load_2d(sa + j * COs * 128, &tm_a, p0, b * p.a_rows_per_batch, &bar_tma[st]);
          tma_load_2d(sb + j * CIs * 128, &tm_b, p0, b * p.b_rows_per_batch, &bar_tma[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA warp =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint32_t sbase = smem_u32(smem);
    for (int n = 0; n < ntile_cta; ++n) {
      const int st = n % NST, use = n / NST;
      mbar_wait(&bar_lo[st], (uint32_t)(use & 1));
      fence_after_sync();
      const uint32_t a0 = sbase + st * stage_bytes;
      uint64_t ah = make_desc_sw128(a0), al = make_desc_sw128(a0 + a_bytes);
      uint64_t bh = make_desc_sw128(a0 + 2 * a_bytes), bl = make_desc_sw128(a0 + 2 * a_bytes + b_bytes);
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {                 // 32 points = four K = 8 steps, 32 bytes apart inside the atom
        const uint32_t acc = (n == 0 && s2 == 0) ? 0u : 1u;
        if (leader) {
          mma_tf32(tmem, ah, bh, idesc, acc);              // hi * hi
          mma_tf32(tmem + N, al, bh, idesc, acc);          // lo * hi
          mma_tf32(tmem + 2 * N, ah, bl, idesc, acc);      // hi * lo
        }
        ah += 2; al += 2; bh += 2; bl += 2;
      }
      if (leader) mma_commit(&bar_empty[st]);
      __syncwarp();
    }
    if (leader) mma_commit(&bar_done);
    __syncwarp();
  } else {
    // ===== builders: tf32 split of the landed tiles, in place =====
    const int bt = tid - 64;
    const int nvecA = NS * COs * 8, nvecB = NS * CIs * 8;          // float4 per tile actually covered by boxes
    for (int n = 0; n < ntile_cta; ++n) {
      const int st = n % NST, use = n / NST;
      mbar_wait(&bar_tma[st], (uint32_t)(use & 1));
      float4* ah = reinterpret_cast<float4*>(smem + st * stage_bytes);
      float4* al = ah + a_bytes / 16;
      float4* bh = al + a_bytes / 16;
      float4* bl = bh + b_bytes / 16;
      const int t = (int)blockIdx.x + n * (int)gridDim.x;
      const int nvalid = min(NS, p.nbatch - (t / p.nchunks) * NS);
      const int validA = nvalid * COs * 8, validB = nvalid * CIs * 8;   // float4 of the slots that were loaded
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = bt; i < nvecA; i += kWtBuilders) {
        float4 h = zero4, l = zero4;
        if (i < validA) {
          const float4 v = ah[i];
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        }
        ah[i] = h;
        al[i] = l;
      }
      for (int i = bt; i < nvecB; i += kWtBuilders) {
        float4 h = zero4, l = zero4;
        if (i < validB) {
          const float4 v = bh[i];
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        }
        bh[i] = h;
        bl[i] = l;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_lo[st]);
    }
  }

  // ---- epilogue: D0 + D1 + D2, diagonal blocks summed over the NS slots, one partial per CTA ----
  if (warp >= 2) {
    const int bt = tid - 64;
    float* stg = reinterpret_cast<float*>(smem);                   // [NS][CO * CI] staging (the stages are free now)
    float* out = p.partial + (long long)blockIdx.x * (CO * CI);
    if (ntile_cta > 0) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
      // tcgen05.ld reaches only the TMEM lane quadrant (warp index % 4) of the issuing warp, whatever the address says:
      // warps 2..9 -> quadrants 2, 3, 0, 1, 2, 3, 0, 1; the first four take the low column half, the last four the high
      const int wq = warp & 3, wh = (warp - 2) >> 2;
      const int row = wq * 32 + lane;
      const int j = row / COs, o = row - j * COs;
      const bool valid = j < NS && o < CO;
      const int nhalf = (N / 16 + 1) / 2;                          // 16-column groups per half
      for (int g = wh * nhalf; g < min((wh + 1) * nhalf, N / 16); ++g) {
        const int c0 = g * 16;
        float v[16];
        tmem_ld16x3_sum(tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, (uint32_t)N, v);
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int i = c0 + jj - j * CIs;
            if (i >= 0 && i < CI) stg[(j * CO + o) * CI + i] = v[jj];
          }
        }
      }
      fence_before_sync();
      nbar_sync(kWtBarEpi, kWtBuilders);
      for (int e = bt; e < CO * CI; e += kWtBuilders) {
        float s = 0.f;
        for (int jj = 0; jj < NS; ++jj) s += stg[jj * CO * CI + e];
        out[e] = s;
      }
    } else {
      for (int e = bt; e < CO * CI; e += kWtBuilders) out[e] = 0.f;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace tc
}  // namespace deno

namespace deno {
namespace tc {

// ----------------------------------------------------------------------------------------------
// plane_analysis_tma: same contract as plane_analysis_tc_kernel - truncated forward rDFT of every (b, c) plane, and
// (FUSE_GZ) gz = gy * act'(z) written once on the way - with the planes brought in by tensor-map TMA.
//
// Rows of a plane are W floats apart, which is a multiple of 16 bytes only every g = 4 / gcd(W, 4) rows, so the
// flat [rows][W] matrix is described by g tensor maps, one per row class j = row mod g (row stride g*W*4 bytes).  A TMA
// box must start on a 16-byte boundary in global memory (an inner coordinate of 2 floats is an illegal instruction on
// hardware, profiles/r2g_tma_box_probe.txt), so class j starts at the 16-byte boundary BELOW its first row: every landed
// row carries shift_j = (j*W) mod 4 leading floats of the previous row, and class j has its own twiddle image E_j whose
// column kc stands for w = kc - shift_j (zero columns under the foreign floats).  A box [H/g rows][32 floats] of class j
// of plane pp lands as rows ((j*P + pp)*HGp + hh) of a K-major SWIZZLE_128B tile: the B operand (per class) of
//   stage 1 (transposed)   D1[n, r] = sum_w E[n, w] * x[r, w]      A = E image: rows n < N1 tf32 hi, rows N1 + n lo
//                          two MMAs per K = 8 step: B = tf32(x) (masked in place) and B = x - tf32(x), one accumulator
//                          -> T[n, r] = D1[n, r] + D1[N1 + n, r]       (n = 2l: Re, 2l+1: Im of sum_w x e^{-i psi_l w})
//   conversion             T -> tf32 hi / lo rows of the stage-2 B image  Timg[(part, pp, n), K' = j*HGp + hh]
//   stage 2                D2[(grp, k), (part, pp, n)] = sum_K' A2[(grp, k), K'] * Timg[(part, pp, n), K']
//                          A2 rows: grp 0 cos hi, 1 sin hi, 2 cos lo, 3 sin lo of phi_k h (h = g*hh + j); ONE MMA per step
//   epilogue               C = D2[cos hi][hi] + D2[cos hi][lo] + D2[cos lo][hi],  S likewise with the sin rows
//                          X^[k, l] = (C[2l] + S[2l+1]) + i (C[2l+1] - S[2l]),  scaled, -> global
// Warp roles: 0-3 conversion + epilogue (TMEM lane quadrants), 4 TMA producer, 5 MMA issuer, 6-13 builders (tf32
// split of the landed tiles; FUSE_GZ: the multiply and the TMA store of gz).  Everything is handed over through
// mbarriers; the tensor pipe order is S1(t), S2(t-1), S1(t+1), ... so the conversion of tile t runs under S1(t+1).
// ----------------------------------------------------------------------------------------------
struct AnalysisTmaParams {
  int H, W, g, HG, HGp, P;       // rows, cols, row classes, rows per class (H / g), padded to 16, planes per tile
  int KH, KHp, MW, N1;           // packed modes along H (padded to 8), modes along W, E rows per part (2 MW padded to 8)
  int C, X, NP, ntiles;          // channels, outer extent (3-D), planes in memory order, tiles
  int nchunks;                   // ceil(W / 32)
  int N, K2p, N2;                // stage-1 N (g*P*HGp), stage-2 K (padded to 32), stage-2 N (2*P*N1)
  int shift[4];                  // leading foreign floats of every landed row of class j (host-side image construction)
  int nsh, nsl;                  // stages of the landed-tile ring / the lo-tile ring
  const float* e1;               // E image, final shared-memory layout: nchunks x [2 N1 rows][32] SWIZZLE_128B
  const float* a2;               // A2 image: (K2p / 32) x [4 KHp rows][32] SWIZZLE_128B
  const float* cl;
  float scale;
  int herm;
  float2* out;                   // [B * X][C][KH][MW]
  const float* u_base;           // x / gy (L2 prefetch only)
  const float* z_base;
  int l2_prefetch;               // plane bytes are 16-byte multiples: bulk L2 prefetch of upcoming tiles
  int dbg;                       // bring-up switches (DENO_AT_DBG): 2 no MMA issue, 4 no TMA loads, 8 no TMEM loads
};

struct AnalysisTmaMaps { CUtensorMap u[4], z[4], gz[4]; };

constexpr int kAtEpiWarps = 4, kAtBuildWarps = 8;
constexpr int kAtThreads = (kAtEpiWarps + 2 + kAtBuildWarps) * 32;
constexpr int kAtBarConv = 4, kAtBarEpi = 5, kAtBarBuild = 6;

struct AnalysisTmaSmem { int e1, a2, timg, hi, lo, stg, xch, total; int hi_stage, lo_stage, e1_bytes, a2_bytes, t_bytes, xch_pitch; };

__host__ __device__ inline AnalysisTmaSmem analysis_tma_smem(const AnalysisTmaParams& p, bool fuse_gz) {
  AnalysisTmaSmem m;
  m.e1_bytes = p.g * p.nchunks * 2 * p.N1 * 128;       // one twiddle image per row class (columns shifted by shift[j])
  m.a2_bytes = (p.K2p / 32) * 4 * p.KHp * 128;
  m.t_bytes = (p.K2p / 32) * p.N2 * 128;
  m.hi_stage = p.N * 128 * (fuse_gz ? 2 : 1);
  m.lo_stage = p.N * 128;
  m.xch_pitch = p.P * p.N1 + 1;
  int off = 0;
  m.e1 = off;   off += (m.e1_bytes + 1023) / 1024 * 1024;
  m.a2 = off;   off += (m.a2_bytes + 1023) / 1024 * 1024;
  m.timg = off; off += (m.t_bytes + 1023) / 1024 * 1024;
  m.hi = off;   off += p.nsh * m.hi_stage;
  m.lo = off;   off += p.nsl * m.lo_stage;
  m.stg = off;  off += 2 * p.N1 * 16 * 4;
  m.xch = off;  off += (2 * p.KHp * m.xch_pitch * 4 + 15) / 16 * 16;
  m.total = off;
  return m;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src_smem)) : "memory");
}

template <bool FUSE_GZ>
__global__ void __launch_bounds__(kAtThreads, 1)
plane_analysis_tma_kernel(const __grid_constant__ AnalysisTmaMaps maps, AnalysisTmaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_hi_full[4], bar_hi_empty[4], bar_lo_full[4], bar_lo_empty[4];
  __shared__ __align__(8) uint64_t bar_d1_full[2], bar_d1_empty[2], bar_t_full, bar_t_empty, bar_d2_full, bar_d2_empty, bar_const;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const AnalysisTmaSmem so = analysis_tma_smem(p, FUSE_GZ);
  const int N = p.N, N1 = p.N1, N2 = p.N2, P = p.P, g = p.g, HG = p.HG, HGp = p.HGp;
  const int NSH = p.nsh, NSL = p.nsl;
  const int nk2 = p.K2p / 32;
  uint32_t ncols = 32;
  while ((int)ncols < 2 * N + N2) ncols *= 2;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 32) {
    for (int s = 0; s < NSH; ++s) { mbar_init(&bar_hi_full[s], 1); mbar_init(&bar_hi_empty[s], FUSE_GZ ? 2 : 1); }
    for (int s = 0; s < NSL; ++s) { mbar_init(&bar_lo_full[s], kAtBuildWarps); mbar_init(&bar_lo_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_d1_full[b], 1); mbar_init(&bar_d1_empty[b], 2); }
    mbar_init(&bar_t_full, 2); mbar_init(&bar_t_empty, 1);
    mbar_init(&bar_d2_full, 1); mbar_init(&bar_d2_empty, kAtEpiWarps);
    mbar_init(&bar_const, 1);
    mbar_fence_init();
    mbar_expect_tx(&bar_const, (uint32_t)(so.e1_bytes + so.a2_bytes));
    bulk_g2s(smem + so.e1, p.e1, (uint32_t)so.e1_bytes, &bar_const);
    bulk_g2s(smem + so.a2, p.a2, (uint32_t)so.a2_bytes, &bar_const);
  }
  // Tile rows no box ever writes (the pad rows hh >= HG of every 16-row-padded block, the K' pad of the stage-2 image)
  // are read by the MMAs and multiplied by zero twiddles: they must be finite, so everything starts as zeros.
  for (int i = tid; i < (so.total - so.timg) / 16; i += kAtThreads)
    reinterpret_cast<float4*>(smem + so.timg)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  DENO_PDL_SYNC();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_d2 = tmem + (uint32_t)(2 * N);
  const int ntile_cta = (p.ntiles > (int)blockIdx.x) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nch = p.nchunks;

  if (warp == 4) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      const uint32_t tx = (uint32_t)(g * P * HG * 128 * (FUSE_GZ ? 2 : 1));
      const long long plane_bytes = (long long)p.H * p.W * 4;
      auto l2_prefetch = [&](int n) {
        if (!p.l2_prefetch || n >= ntile_cta) return;
        const long long m0 = ((long long)blockIdx.x + (long long)n * gridDim.x) * P;
        const int np = (int)min((long long)P, (long long)p.NP - m0);
        if (np <= 0) return;
        const uint32_t bytes = (uint32_t)(np * plane_bytes);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.u_base + m0 * p.H * p.W), "r"(bytes) : "memory");
        if (FUSE_GZ) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.z_base + m0 * p.H * p.W), "r"(bytes) : "memory");
      };
      l2_prefetch(0); l2_prefetch(1); l2_prefetch(2);
      int ci = 0;
      for (int n = 0; n < ntile_cta; ++n) {
        l2_prefetch(n + 3);
        const int m0 = ((int)blockIdx.x + n * (int)gridDim.x) * P;
        for (int c = 0; c < nch; ++c, ++ci) {
          const int st = ci % NSH, use = ci / NSH;
          mbar_wait(&bar_hi_empty[st], (uint32_t)((use & 1) ^ 1));
          unsigned char* dst = smem + so.hi + st * so.hi_stage;
          if (p.dbg & 4) { mbar_arrive(&bar_hi_full[st]); continue; }
          mbar_expect_tx(&bar_hi_full[st], tx);
          for (int j = 0; j < g; ++j)
            for (int pp = 0; pp < P; ++pp) {
              const int row0 = (j * P + pp) * HGp;
              tma_load_2d(dst + row0 * 128, &maps.u[j], 32 * c, (m0 + pp) * HG, &bar_hi_full[st]);
              if (FUSE_GZ) tma_load_2d(dst + N * 128 + row0 * 128, &maps.z[j], 32 * c, (m0 + pp) * HG, &bar_hi_full[st]);
            }
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================================== MMA issuer
    const bool leader = elect_one();
    const bool issue = leader && !(p.dbg & 2);
    const int NC = P * HGp;                            // stage-1 N of one row class
    const uint32_t idesc1 = make_idesc_tf32(128, NC), idesc2 = make_idesc_tf32(128, N2);
    const uint32_t sbase = smem_u32(smem);
    mbar_wait(&bar_const, 0);
    auto stage2 = [&](int n) {                         // S2 of this CTA's n-th tile
      mbar_wait(&bar_t_full, (uint32_t)(n & 1));
      mbar_wait(&bar_d2_empty, (uint32_t)((n & 1) ^ 1));
      fence_after_sync();
      for (int c = 0; c < nk2; ++c) {
        uint64_t a = make_desc_sw128(sbase + so.a2 + c * (4 * p.KHp * 128));
        uint64_t b = make_desc_sw128(sbase + so.timg + c * (N2 * 128));
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          if (issue) mma_tf32(tmem_d2, a, b, idesc2, (c | s2) ? 1u : 0u);
          a += 2; b += 2;
        }
      }
      if (leader) { mma_commit(&bar_t_empty); mma_commit(&bar_d2_full); }
      __syncwarp();
    };
    int ci = 0;
    for (int n = 0; n < ntile_cta; ++n) {
      const int b = n & 1;
      mbar_wait(&bar_d1_empty[b], (uint32_t)(((n >> 1) & 1) ^ 1));
      fence_after_sync();
      const uint32_t d1 = tmem + (uint32_t)(b * N);
      for (int c = 0; c < nch; ++c, ++ci) {
        const int sl = ci % NSL, sh = ci % NSH;
        mbar_wait(&bar_lo_full[sl], (uint32_t)((ci / NSL) & 1));      // implies the landed tile is complete (and masked)
        fence_after_sync();
        for (int j = 0; j < g; ++j) {                    // per row class: its own (shifted) twiddle image, its rows of the tile
          uint64_t e = make_desc_sw128(sbase + so.e1 + (j * nch + c) * (2 * N1 * 128));
          uint64_t xh = make_desc_sw128(sbase + so.hi + sh * so.hi_stage + j * NC * 128);
          uint64_t xl = make_desc_sw128(sbase + so.lo + sl * so.lo_stage + j * NC * 128);
          const uint32_t dj = d1 + (uint32_t)(j * NC);
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2) {
            if (issue) {
              mma_tf32(dj, e, xh, idesc1, (c | s2) ? 1u : 0u);
              mma_tf32(dj, e, xl, idesc1, 1u);
            }
            e += 2; xh += 2; xl += 2;
          }
        }
        if (leader) { mma_commit(&bar_hi_empty[sh]); mma_commit(&bar_lo_empty[sl]); }
        __syncwarp();
      }
      if (leader) mma_commit(&bar_d1_full[b]);
      __syncwarp();
      if (n > 0) stage2(n - 1);
    }
    if (ntile_cta > 0) stage2(ntile_cta - 1);
  } else if (warp >= 6) {
    // ===================================================================== builders
    const int bt = tid - 6 * 32, nb = kAtBuildWarps * 32;
    const int nvec = N * 8;                                    // float4 per tile
    int ci = 0;
    for (int n = 0; n < ntile_cta; ++n) {
      const int m0 = ((int)blockIdx.x + n * (int)gridDim.x) * P;
      for (int c = 0; c < nch; ++c, ++ci) {
        const int sh = ci % NSH, sl = ci % NSL;
        mbar_wait(&bar_hi_full[sh], (uint32_t)((ci / NSH) & 1));
        mbar_wait(&bar_lo_empty[sl], (uint32_t)(((ci / NSL) & 1) ^ 1));
        float4* hi = reinterpret_cast<float4*>(smem + so.hi + sh * so.hi_stage);
        float4* zt = hi + nvec;
        float4* lo = reinterpret_cast<float4*>(smem + so.lo + sl * so.lo_stage);
        for (int i = bt; i < nvec; i += nb) {
          float4 v = hi[i];
          if (FUSE_GZ) {
            const float4 z = zt[i];
            v = make_float4(v.x * z.x, v.y * z.y, v.z * z.z, v.w * z.w);
            zt[i] = v;                                           // gz itself: stored to global from here by TMA
          }
          float4 h, l;
          split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
          hi[i] = h;
          lo[i] = l;
        }
        fence_proxy_async();
        if (FUSE_GZ) {
          nbar_sync(kAtBarBuild, nb);                            // the whole gz tile is in shared memory
          if (bt == 0) {
            for (int j = 0; j < g; ++j)
              for (int pp = 0; pp < P; ++pp)
                if (m0 + pp < p.NP)
                  tma_store_2d(&maps.gz[j], reinterpret_cast<unsigned char*>(zt) + (j * P + pp) * HGp * 128,
                               32 * c, (m0 + pp) * HG);   // the shift[j] leading floats are the previous row's tail: same values its own class writes
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_lo_full[sl]);
        if (FUSE_GZ && bt == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the store has read the tile: stage reusable
          mbar_arrive(&bar_hi_empty[sh]);
        }
      }
    }
    if (FUSE_GZ && bt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // gz fully written before exit
  } else {
    // ===================================================================== conversion + epilogue (warps 0-3)
    const int L = warp * 32 + lane;                    // TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float* stg = reinterpret_cast<float*>(smem + so.stg);       // [2][N1][16]
    float* xch = reinterpret_cast<float*>(smem + so.xch);       // [2][KHp][xch_pitch]
    const int XP = so.xch_pitch;
    const int PN = P * N1;
    mbar_wait(&bar_const, 0);
    auto conversion = [&](int n) {
      if (warp >= 2) return;
      const int b = n & 1;
      mbar_wait(&bar_d1_full[b], (uint32_t)((n >> 1) & 1));
      mbar_wait(&bar_t_empty, (uint32_t)((n & 1) ^ 1));
      fence_after_sync();
      const uint32_t d1 = tmem + (uint32_t)(b * N) + lane_addr;
      const bool is_lo = L >= N1 && L < 2 * N1, is_hi = L < N1;
      const int nn = is_lo ? L - N1 : L;
      int sb = 0;
      for (int j = 0; j < g; ++j)
        for (int pp = 0; pp < P; ++pp)
          for (int u = 0; u < HGp / 16; ++u) {
            float v[16];
            if (p.dbg & 8) { for (int q = 0; q < 16; ++q) v[q] = 0.f; }
            else tmem_ld16(d1 + (uint32_t)((j * P + pp) * HGp + 16 * u), v);
            float4* sg = reinterpret_cast<float4*>(stg + (sb * N1 + nn) * 16);
            if (is_lo) {
#pragma unroll
              for (int q = 0; q < 4; ++q) sg[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            nbar_sync(kAtBarConv, 64);
            if (is_hi) {
              const int kp = j * HGp + 16 * u;                   // K' of the first of the 16 values
              const int chunk = kp >> 5, c16 = (kp & 31) >> 2;
              const int row_hi = pp * N1 + nn, row_lo = PN + row_hi;
              unsigned char* base = smem + so.timg + chunk * (N2 * 128);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 s4 = sg[q];
                float4 h, l;
                split_tf32(v[4 * q] + s4.x, h.x, l.x);
                split_tf32(v[4 * q + 1] + s4.y, h.y, l.y);
                split_tf32(v[4 * q + 2] + s4.z, h.z, l.z);
                split_tf32(v[4 * q + 3] + s4.w, h.w, l.w);
                *reinterpret_cast<float4*>(base + row_hi * 128 + (((c16 + q) ^ (row_hi & 7)) << 4)) = h;
                *reinterpret_cast<float4*>(base + row_lo * 128 + (((c16 + q) ^ (row_lo & 7)) << 4)) = l;
              }
            }
            sb ^= 1;
          }
      fence_before_sync();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bar_d1_empty[b]); mbar_arrive(&bar_t_full); }
    };
    auto epilogue = [&](int n) {
      const int m0 = ((int)blockIdx.x + n * (int)gridDim.x) * P;
      mbar_wait(&bar_d2_full, (uint32_t)(n & 1));
      fence_after_sync();
      const int grp = L / p.KHp, k = L - grp * p.KHp;
      const bool row_ok = grp < 4 && k < p.KH;
      const uint32_t d2 = tmem_d2 + lane_addr;
      // phase 1: the tf32-hi twiddle rows (cos: grp 0, sin: grp 1) write  D2[.., hi part] + D2[.., lo part]
      // phase 2: the lo twiddle rows (grp 2, 3) add  D2[.., hi part]
      for (int ph = 0; ph < 2; ++ph) {
        const bool mine = row_ok && ((ph == 0) ? grp < 2 : grp >= 2);
        if (warp * 32 < 4 * p.KHp) {                   // warp holds twiddle rows at all (tcgen05.ld is warp-collective)
          for (int c0 = 0; c0 < PN; c0 += 16) {
            if ((ph == 0 && warp * 32 >= 2 * p.KHp) || (ph == 1 && warp * 32 + 32 <= 2 * p.KHp)) continue;
            float a[16], bb[16];
            if (p.dbg & 8) { for (int q = 0; q < 16; ++q) { a[q] = 0.f; bb[q] = 0.f; } }
            else {
              tmem_ld16(d2 + (uint32_t)c0, a);
              if (ph == 0) tmem_ld16(d2 + (uint32_t)(PN + c0), bb);
            }
            if (mine) {
              float* dst = xch + ((grp & 1) * p.KHp + k) * XP + c0;
#pragma unroll
              for (int q = 0; q < 16; ++q)
                if (c0 + q < PN) dst[q] = (ph == 0) ? a[q] + bb[q] : dst[q] + a[q];
            }
          }
        }
        if (ph == 1) {
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_d2_empty);
        }
        nbar_sync(kAtBarEpi, kAtEpiWarps * 32);
      }
      // combine: one thread per (plane, k, l)
      const int nitem = P * p.KH * p.MW;
      for (int it = L; it < nitem; it += kAtEpiWarps * 32) {
        const int l = it % p.MW, t2 = it / p.MW;
        const int k2 = t2 % p.KH, pp = t2 / p.KH;
        const int m = m0 + pp;
        if (m < p.NP) {
          const int CX = p.C * p.X;
          const int bo = m / CX, r = m - bo * CX, c = r / p.X, xi = r - c * p.X;
          const long long q_out = ((long long)bo * p.X + xi) * p.C + c;
          const float* cr = xch + k2 * XP + pp * N1 + 2 * l;
          const float* sr = xch + (p.KHp + k2) * XP + pp * N1 + 2 * l;
          const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
          p.out[q_out * p.KH * p.MW + (long long)k2 * p.MW + l] = make_float2((cr[0] + sr[1]) * f, (cr[1] - sr[0]) * f);
        }
      }
      nbar_sync(kAtBarEpi, kAtEpiWarps * 32);          // xch is rewritten by the next tile's phase 1
    };
    for (int n = 0; n < ntile_cta; ++n) {
      conversion(n);
      if (n > 0) epilogue(n - 1);
    }
    if (ntile_cta > 0) epilogue(ntile_cta - 1);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace tc
}  // namespace deno
