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
      const uint32_t tx = (uint32_t)(NS * COs * 128 + NS * CIs * 128);
      for (int n = 0; n < ntile_cta; ++n) {
        const int st = n % NST, use = n / NST;
        mbar_wait(&bar_empty[st], (uint32_t)((use & 1) ^ 1));     // MMAs of the previous use of this stage have retired
        const int t = (int)blockIdx.x + n * (int)gridDim.x;
        const int grp = t / p.nchunks, chunk = t - grp * p.nchunks;
        const int p0 = chunk * 32;
        unsigned char* sa = smem + st * stage_bytes;
        unsigned char* sb = sa + 2 * a_bytes;
        mbar_expect_tx(&bar_tma[st], tx);
        for (int j = 0; j < NS; ++j) {
          const int b = grp * NS + j;                              // beyond the batch: the box is out of bounds -> zeros
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
      for (int i = bt; i < nvecA; i += kWtBuilders) {
        const float4 v = ah[i];
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
        ah[i] = h;
        al[i] = l;
      }
      for (int i = bt; i < nvecB; i += kWtBuilders) {
        const float4 v = bh[i];
        float4 h, l;
        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
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
      const int ew = warp - 2;                                     // 0..7
      const int wq = ew & 3, wh = ew >> 2;                         // TMEM lane quadrant, column half
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
