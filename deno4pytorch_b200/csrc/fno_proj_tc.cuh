// fno_proj_tc.cuh - the projection end of an FNO stack on the tensor cores (tcgen05 / TMEM), for the widths
// the reference's demos use (C = 32 or 64 channels, 128 hidden units).  Same contract as fno_proj_fwd_kernel /
// fno_proj_bwd_kernel (fno_pointwise.cuh), which remain the path for every other shape.
//
// A tile is 128 consecutive padded grid points of one sample = the 128 TMEM lanes.
//   GEMM 1   D1[pt, j] = sum_c h[pt, c] w1[j, c]            A = h tile as a K-major image (K = C), B = w1 as stored
//   forward  y[pt, o]  = b2[o] + sum_j w2[o, j] gelu(D1[pt, j] + b1[j])              (epilogue, registers only)
//   backward gt[pt, j] = (sum_o gy[pt, o] w2[o, j]) gelu'(D1[pt, j] + b1[j])         (epilogue; also stored for gW1)
//   GEMM 2   D2[pt, c] = sum_j gt[pt, j] w1[j, c]            A = gt straight from the epilogue registers: a thread
//            owns one point and consecutive hidden units, i.e. exactly the 16-byte K-major atoms of the A image
// 3xTF32 (hi/lo split of both operands, three products) keeps fp32 accuracy.  One persistent CTA per SM:
// 8 worker warps (lane quadrant x column half) and one MMA-issuing warp; hand-offs to the MMA warp are
// bar.arrive / bar.sync named barriers, completions come back through mbarriers (tcgen05.commit).
// The backward streams GEMM 2 in chunks of 32 hidden units through a two-slot ring so the tensor core runs
// while the workers evaluate the next chunk's GELUs; bias / fc2-weight gradients accumulate in registers over
// all tiles of the CTA and are reduced once at the end (fixed order: deterministic).
#pragma once
#include "fno_pointwise.cuh"
#include "spectral_tc.cuh"

namespace deno {
namespace tc {

constexpr int kPjWorkers = 256;               // backward: 8 worker warps (lane quadrant x column half); 128 accumulators per thread
constexpr int kPjThreads = kPjWorkers + 32;
constexpr int kPjFwdWorkers = 512;            // forward: 16 worker warps (lane quadrant x column quarter)
constexpr int kPjFwdThreads = kPjFwdWorkers + 32;
constexpr int kPjBarA1 = 2;        // A1 image of the next tile stored (workers arrive, MMA warp syncs)
constexpr int kPjBarChunk0 = 3;    // 3, 4: ring slot stored
constexpr int kPjBarWorkers = 5;   // workers only

struct ProjTcParams {
  PointGeom g;
  int B, O;
  const float* h;
  const float* w1;
  const float* b1;
  const float* w2;
  const float* b2;
  float* y;
  const float* gy;
  float* gh;
  float* gt;             // backward: [B * tiles_per_sample][HD][128]
  float* partial;        // backward: [gridDim.x][HD + O * HD + O]
  int tiles_per_sample;  // ceil(nppts / 128)
  int ntiles;
};

template <int C, int HD>
struct ProjTcSmem {
  static constexpr int a_lbo = 16 * 128 + 16;                 // A images: 128 rows, padded K-quad stride
  static constexpr int b1_one = (C / 4) * 16 * HD;            // w1 image, N = HD rows, K = C
  static constexpr int a1_one = (C / 4) * a_lbo;
  static constexpr int b2_lbo = 16 * 2 * C;                   // w1^T image, rows c (hi) then C + c (lo), K = HD
  static constexpr int b2_bytes = (HD / 4) * b2_lbo;
  static constexpr int ring_one = 8 * a_lbo;                  // one hi or lo image of a 32-unit chunk (8 K-quads)
  static constexpr int off_b1 = 0;
  static constexpr int off_a1 = off_b1 + 2 * b1_one;
  static constexpr int off_small = off_a1 + 2 * a1_one;       // b1s [HD], w2s [4][HD], b2s [4], ys [2][4][128][4]
  static constexpr int small_bytes = (HD + 4 * HD + 4 + 2 * 4 * 128 * 4) * 4;
  static constexpr int fwd_total = off_small + small_bytes + 4096;
  static constexpr int off_b2 = off_small + small_bytes;
  static constexpr int off_ring = off_b2 + b2_bytes;
  static constexpr int off_red = off_ring + 4 * ring_one;     // [4][2 * HD + 1] end-of-kernel reduction
  static constexpr int bwd_total = off_red + 4 * (2 * HD + 4) * 4 + 4096;
};

// h tile -> registers: thread (point xrow, column group wg of NG) owns channels [wg * C/NG, (wg+1) * C/NG)
template <int C, int NG>
__device__ __forceinline__ void pj_load_tile(const ProjTcParams& p, int tile, int xrow, int wg, float (&hr)[C / NG], bool& inside,
                                             bool& valid, int& b, int& pp, int& pc) {
  // `tile` is passed as (sample, tile inside the sample) packed by PjTileIter: no division here
  b = tile >> 12;
  pp = (tile & 4095) * 128 + xrow;
  inside = pp < p.g.nppts;
  pc = 0;
  valid = inside && padded_to_cropped(p.g, pp, pc);
  const float* src = p.h + ((long long)b * C + wg * (C / NG)) * p.g.nppts + pp;
#pragma unroll
  for (int i = 0; i < C / NG; ++i) hr[i] = valid ? __ldg(src + (long long)i * p.g.nppts) : 0.f;
}

// The tiles of a CTA are blockIdx.x + k * gridDim.x; (sample, tile inside the sample) advance by addition and compare
// instead of a division per tile and thread.  Packed as (sample << 12) | tile_in_sample (tiles_per_sample <= 4096,
// checked on the host).
struct PjTileIter {
  int b, t, step, tps;
  __device__ __forceinline__ PjTileIter(int first, int step_, int tps_) : b(first / tps_), t(first % tps_), step(step_), tps(tps_) {}
  __device__ __forceinline__ int packed() const { return (b << 12) | t; }
  __device__ __forceinline__ void next() {
    t += step;
    while (t >= tps) { t -= tps; ++b; }
  }
};

template <int C, int HD, int NG>
__device__ __forceinline__ void pj_store_a1(unsigned char* smem, int xrow, int wg, const float (&hr)[C / NG]) {
  using S = ProjTcSmem<C, HD>;
  unsigned char* hi_img = smem + S::off_a1;
  unsigned char* lo_img = hi_img + S::a1_one;
  const int row_off = (xrow & 7) * 16 + (xrow >> 3) * 128;
#pragma unroll
  for (int q = 0; q < C / (4 * NG); ++q) {
    float4 hi, lo;
    split_tf32(hr[4 * q], hi.x, lo.x);
    split_tf32(hr[4 * q + 1], hi.y, lo.y);
    split_tf32(hr[4 * q + 2], hi.z, lo.z);
    split_tf32(hr[4 * q + 3], hi.w, lo.w);
    const int off = row_off + (wg * (C / (4 * NG)) + q) * S::a_lbo;
    *reinterpret_cast<float4*>(hi_img + off) = hi;
    *reinterpret_cast<float4*>(lo_img + off) = lo;
  }
}

// w1 (HD, C) row-major -> B operand images (N = HD rows, K = C), hi / lo; small vectors to shared memory
template <int C, int HD>
__device__ __forceinline__ void pj_prologue(const ProjTcParams& p, unsigned char* smem, int tid) {
  using S = ProjTcSmem<C, HD>;
  const int nthr = (int)blockDim.x;
  const float4* w4 = reinterpret_cast<const float4*>(p.w1);
  for (int i0 = tid; i0 < HD * (C / 4); i0 += 4 * nthr) {         // four loads in flight per thread
    float4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) w[u] = __ldg(w4 + min(i0 + u * nthr, HD * (C / 4) - 1));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < HD * (C / 4)) {
        const int j = i / (C / 4), q = i - j * (C / 4);
        float4 hi, lo;
        split_tf32(w[u].x, hi.x, lo.x);
        split_tf32(w[u].y, hi.y, lo.y);
        split_tf32(w[u].z, hi.z, lo.z);
        split_tf32(w[u].w, hi.w, lo.w);
        const int off = (j & 7) * 16 + (j >> 3) * 128 + q * (16 * HD);
        *reinterpret_cast<float4*>(smem + S::off_b1 + off) = hi;
        *reinterpret_cast<float4*>(smem + S::off_b1 + S::b1_one + off) = lo;
      }
    }
  }
  float* b1s = reinterpret_cast<float*>(smem + S::off_small);
  float* w2s = b1s + HD;
  float* b2s = w2s + 4 * HD;
  for (int i = tid; i < HD; i += (int)blockDim.x) b1s[i] = p.b1[i];
  for (int i = tid; i < p.O * HD; i += (int)blockDim.x) w2s[i] = p.w2[i];
  for (int i = tid; i < 4; i += (int)blockDim.x) b2s[i] = (p.b2 != nullptr && i < p.O) ? p.b2[i] : 0.f;
}

// GEMM 1 of one tile: 3 products per K step into one accumulator
template <int C, int HD>
__device__ __forceinline__ void pj_issue_gemm1(uint32_t sbase, uint32_t tmem_d, bool leader) {
  using S = ProjTcSmem<C, HD>;
  uint64_t ah = make_desc(sbase + S::off_a1, S::a_lbo, 128), al = make_desc(sbase + S::off_a1 + S::a1_one, S::a_lbo, 128);
  uint64_t bh = make_desc(sbase + S::off_b1, 16 * HD, 128), bl = make_desc(sbase + S::off_b1 + S::b1_one, 16 * HD, 128);
  const uint64_t inc_a = (uint64_t)((2 * S::a_lbo) >> 4), inc_b = (uint64_t)((2 * 16 * HD) >> 4);
  const uint32_t idesc = make_idesc_tf32(128, HD);
#pragma unroll
  for (int s = 0; s < C / 8; ++s) {
    if (leader) {
      mma_tf32(tmem_d, al, bh, idesc, s > 0 ? 1u : 0u);   // small terms first
      mma_tf32(tmem_d, ah, bl, idesc, 1u);
      mma_tf32(tmem_d, ah, bh, idesc, 1u);
    }
    ah += inc_a; al += inc_a; bh += inc_b; bl += inc_b;
  }
}

template <int C, int HD>
__global__ void __launch_bounds__(kPjFwdThreads, 1) fno_proj_fwd_tc_kernel(ProjTcParams p) {
  using S = ProjTcSmem<C, HD>;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar_d1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t tmem_cols = 2 * HD <= 256 ? 256u : 512u;
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 32) {
    mbar_init(&mbar_d1, 1);
    mbar_fence_init();
  }
  pj_prologue<C, HD>(p, smem, tid);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  DENO_PDL_SYNC();   // everything above read weights only; h (and gy) come from the predecessor
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int nmine = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kPjFwdWorkers / 32) {
    const bool leader = elect_one();
    for (int k = 0; k < nmine; ++k) {
      nbar_sync(kPjBarA1, kPjFwdThreads);
      fence_after_sync();
      pj_issue_gemm1<C, HD>(sbase, tmem + (uint32_t)((k & 1) * HD), leader);
      if (leader) mma_commit(&mbar_d1);
      __syncwarp();
    }
  } else {
    const int wq = warp & 3, wg = warp >> 2;
    const int xrow = wq * 32 + lane;
    const float* b1s = reinterpret_cast<const float*>(smem + S::off_small);
    const float* w2s = b1s + HD;
    const float* b2s = w2s + 4 * HD;
    float* ys = const_cast<float*>(b2s) + 4;                       // [2 (tile parity)][4 (wg)][128][4]
    float hr[C / 4];
    bool inside, valid, valid_cur = false;
    int b, pp, pc, b_cur = 0, pc_cur = 0;
    uint32_t phase = 0;
    PjTileIter it((int)blockIdx.x, (int)gridDim.x, p.tiles_per_sample);   // tile of the NEXT load
    if (nmine > 0) {
      pj_load_tile<C, 4>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
      it.next();
      pj_store_a1<C, HD, 4>(smem, xrow, wg, hr);
      fence_proxy_async();
      nbar_arrive(kPjBarA1, kPjFwdThreads);
      valid_cur = valid; b_cur = b; pc_cur = pc;
      if (nmine > 1) {
        pj_load_tile<C, 4>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
        it.next();
      }
    }
    for (int k = 0; k < nmine; ++k) {
      const long long t0 = DENO_CLK();
      mbar_wait(&mbar_d1, phase);
      fence_after_sync();
      phase ^= 1;
      const long long t1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(16, t1 - t0);          // slot 16: wait for GEMM 1
      const bool v_k = valid_cur;
      const int b_k = b_cur, pc_k = pc_cur;
      if (k + 1 < nmine) {                               // GEMM 1 (k) is done with the A1 image: next tile in
        pj_store_a1<C, HD, 4>(smem, xrow, wg, hr);
        fence_proxy_async();
        nbar_arrive(kPjBarA1, kPjFwdThreads);
        valid_cur = valid; b_cur = b; pc_cur = pc;
        if (k + 2 < nmine) {
          pj_load_tile<C, 4>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
          it.next();
        }
      }
      const long long t2 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(17, t2 - t1);          // slot 17: A1 image of the next tile + prefetch issue
      float yo[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t td = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)((k & 1) * HD + wg * (HD / 4));
#pragma unroll
      for (int c0 = 0; c0 < HD / 4; c0 += 16) {
        float v[16];
        tmem_ld16(td + c0, v);
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {            // two hidden units per packed fp32x2 evaluation
          const int j = wg * (HD / 4) + c0 + jj;
          const float2 bb = *reinterpret_cast<const float2*>(b1s + j);
          float2 a, da;
          gelu_fast2(__fadd2_rn(make_float2(v[jj], v[jj + 1]), bb), a, da);
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < p.O) {
              const float2 ww = *reinterpret_cast<const float2*>(w2s + o * HD + j);
              yo[o] = fmaf(ww.y, a.y, fmaf(ww.x, a.x, yo[o]));
            }
        }
      }
      fence_before_sync();                               // D1 reads ordered before the next hand-off
      const long long t3 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(18, t3 - t2);          // slot 18: epilogue math
      float* ysk = ys + (k & 1) * (4 * 128 * 4);
      *reinterpret_cast<float4*>(ysk + (wg * 128 + xrow) * 4) = make_float4(yo[0], yo[1], yo[2], yo[3]);
      nbar_sync(kPjBarWorkers, kPjFwdWorkers);
      if (wg == 0 && v_k) {
        const float4 o0 = *reinterpret_cast<const float4*>(ysk + xrow * 4);
        const float4 o1 = *reinterpret_cast<const float4*>(ysk + (128 + xrow) * 4);
        const float4 o2 = *reinterpret_cast<const float4*>(ysk + (256 + xrow) * 4);
        const float4 o3 = *reinterpret_cast<const float4*>(ysk + (384 + xrow) * 4);
        const float r[4] = {(o0.x + o1.x) + (o2.x + o3.x), (o0.y + o1.y) + (o2.y + o3.y), (o0.z + o1.z) + (o2.z + o3.z),
                            (o0.w + o1.w) + (o2.w + o3.w)};
        float* dst = p.y + ((long long)b_k * p.g.npts + pc_k) * p.O;
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < p.O) dst[o] = r[o] + b2s[o];
      }
      if (warp == 1) { DENO_DBG_ADD(19, DENO_CLK() - t3); DENO_DBG_ADD(20, 1); }   // slots 19/20: exchange + store, tiles
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

// Backward.  O == 1 (the fc2-weight gradient accumulators live in registers).
template <int C, int HD>
__global__ void __launch_bounds__(kPjThreads, 1) fno_proj_bwd_tc_kernel(ProjTcParams p) {
  using S = ProjTcSmem<C, HD>;
  constexpr int NCH = HD / 32;                           // ring chunks per tile: 16 hidden units per column half
  static_assert(NCH >= 2 && NCH % 2 == 0, "the two-slot ring bookkeeping assumes an even number of chunks");
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar_d1, mbar_slot[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t tmem_cols = HD + 3 * C <= 256 ? 256u : 512u;
  if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
  if (tid == 32) {
    mbar_init(&mbar_d1, 1);
    mbar_init(&mbar_slot[0], 1);
    mbar_init(&mbar_slot[1], 1);
    mbar_fence_init();
  }
  pj_prologue<C, HD>(p, smem, tid);
  // w1^T as the B operand of GEMM 2: rows c (hi) and C + c (lo), K = hidden unit j; 16-byte loads, four in flight
  {
    const float4* w4 = reinterpret_cast<const float4*>(p.w1);
    for (int i0 = tid; i0 < HD * (C / 4); i0 += 4 * kPjThreads) {
      float4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = __ldg(w4 + min(i0 + u * kPjThreads, HD * (C / 4) - 1));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kPjThreads;
        if (i < HD * (C / 4)) {
          const int j = i / (C / 4), c0 = (i - j * (C / 4)) * 4;
          const int koff = (j & 3) * 4 + (j >> 2) * S::b2_lbo;
          const float wv[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            float hi, lo;
            split_tf32(wv[e], hi, lo);
            *reinterpret_cast<float*>(smem + S::off_b2 + (c & 7) * 16 + (c >> 3) * 128 + koff) = hi;
            *reinterpret_cast<float*>(smem + S::off_b2 + ((C + c) & 7) * 16 + ((C + c) >> 3) * 128 + koff) = lo;
          }
        }
      }
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  DENO_PDL_SYNC();   // everything above read weights only; h and gy come from the predecessor
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + HD;    // D2: [hi*hi | hi*lo | lo*hi], C columns each
  const uint32_t sbase = smem_u32(smem);
  const int nmine = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kPjWorkers / 32) {
    const bool leader = elect_one();
    const uint32_t idesc_n2 = make_idesc_tf32(128, 2 * C), idesc_n1 = make_idesc_tf32(128, C);
    if (nmine > 0) {
      nbar_sync(kPjBarA1, kPjThreads);
      fence_after_sync();
      pj_issue_gemm1<C, HD>(sbase, tmem_d1, leader);
      if (leader) mma_commit(&mbar_d1);
      __syncwarp();
    }
    for (int k = 0; k < nmine; ++k) {
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int slot = c & 1;
        nbar_sync(kPjBarChunk0 + slot, kPjThreads);
        fence_after_sync();
        const uint32_t ring = sbase + S::off_ring + slot * 2 * S::ring_one;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {                  // two K steps of 8 hidden units per column half
            const int j0 = half * (HD / 2) + c * 16 + s * 8;
            const uint64_t ah = make_desc(ring + (half * 4 + s * 2) * S::a_lbo, S::a_lbo, 128);
            const uint64_t al = make_desc(ring + S::ring_one + (half * 4 + s * 2) * S::a_lbo, S::a_lbo, 128);
            const uint64_t bb = make_desc(sbase + S::off_b2 + (j0 / 4) * S::b2_lbo, S::b2_lbo, 128);
            const uint32_t acc = (c > 0 || half > 0 || s > 0) ? 1u : 0u;
            if (leader) {
              mma_tf32(tmem_d2, ah, bb, idesc_n2, acc);             // hi * [hi | lo]
              mma_tf32(tmem_d2 + 2 * C, al, bb, idesc_n1, acc);     // lo * hi
            }
          }
        }
        if (leader) mma_commit(&mbar_slot[slot]);
        __syncwarp();
      }
      if (k + 1 < nmine) {
        nbar_sync(kPjBarA1, kPjThreads);
        fence_after_sync();
        pj_issue_gemm1<C, HD>(sbase, tmem_d1, leader);
        if (leader) mma_commit(&mbar_d1);
        __syncwarp();
      }
    }
  } else {
    const int wq = warp & 3, wg = warp >> 2;
    const int xrow = wq * 32 + lane;
    const float* b1s = reinterpret_cast<const float*>(smem + S::off_small);
    const float* w2s = b1s + HD;
    float hr[C / 2];
    float acc_b1[HD / 2], acc_w2[HD / 2], acc_b2 = 0.f;
#pragma unroll
    for (int i = 0; i < HD / 2; ++i) { acc_b1[i] = 0.f; acc_w2[i] = 0.f; }
    bool inside, valid, inside_cur = false;
    int b, pp, pc, b_cur = 0, pp_cur = 0;
    float gy_next = 0.f, gy_cur = 0.f;
    uint32_t phase = 0, slot_phase[2] = {0u, 0u};
    PjTileIter it((int)blockIdx.x, (int)gridDim.x, p.tiles_per_sample);   // tile of the NEXT load
    if (nmine > 0) {
      pj_load_tile<C, 2>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
      it.next();
      gy_cur = valid ? __ldg(p.gy + (long long)b * p.g.npts + pc) : 0.f;
      pj_store_a1<C, HD, 2>(smem, xrow, wg, hr);
      fence_proxy_async();
      nbar_arrive(kPjBarA1, kPjThreads);
      inside_cur = inside; b_cur = b; pp_cur = pp;
      if (nmine > 1) {
        pj_load_tile<C, 2>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
        it.next();
        gy_next = valid ? __ldg(p.gy + (long long)b * p.g.npts + pc) : 0.f;
      }
    }
    const int row_off = (xrow & 7) * 16 + (xrow >> 3) * 128;
    for (int k = 0; k < nmine; ++k) {
      const long long t0 = DENO_CLK();
      mbar_wait(&mbar_d1, phase);
      fence_after_sync();
      phase ^= 1;
      const long long t1 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(21, t1 - t0);          // slot 21: wait for GEMM 1
      long long t_slot = 0;
      // gt in blocks of 128 points: [tile][HD][128], every tile one contiguous 64 KB run (bypass_wgrad_tc's gz_tps mode)
      float* gt_dst = p.gt + ((long long)(blockIdx.x + k * gridDim.x) * HD + wg * (HD / 2)) * 128 + xrow;
      if (wg == 0) acc_b2 += gy_cur;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int slot = c & 1;
        float v[16];
        tmem_ld16(tmem_d1 + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wg * (HD / 2) + c * 16), v);
        const float2 gy2 = make_float2(gy_cur, gy_cur);
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {            // two hidden units per packed fp32x2 evaluation
          const int j = wg * (HD / 2) + c * 16 + jj;
          const float2 bb = *reinterpret_cast<const float2*>(b1s + j);
          const float2 ww = *reinterpret_cast<const float2*>(w2s + j);
          float2 a, da;
          gelu_fast2(__fadd2_rn(make_float2(v[jj], v[jj + 1]), bb), a, da);
          const float2 gt = __fmul2_rn(__fmul2_rn(gy2, ww), da);   // zero for padding lanes: their gy is zero
          acc_b1[c * 16 + jj] += gt.x;
          acc_b1[c * 16 + jj + 1] += gt.y;
          acc_w2[c * 16 + jj] = fmaf(gy_cur, a.x, acc_w2[c * 16 + jj]);
          acc_w2[c * 16 + jj + 1] = fmaf(gy_cur, a.y, acc_w2[c * 16 + jj + 1]);
          v[jj] = gt.x;
          v[jj + 1] = gt.y;
          gt_dst[(c * 16 + jj) * 128] = gt.x;
          gt_dst[(c * 16 + jj + 1) * 128] = gt.y;
        }
        if (c >= 2) {                                    // the slot's previous chunk has been consumed by the tensor core
          const long long ts = DENO_CLK();
          mbar_wait(&mbar_slot[slot], slot_phase[slot]);
          slot_phase[slot] ^= 1;
          t_slot += DENO_CLK() - ts;
        }
        unsigned char* ring_hi = smem + S::off_ring + slot * 2 * S::ring_one + (wg * 4) * S::a_lbo + row_off;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 hi, lo;
          split_tf32(v[4 * q], hi.x, lo.x);
          split_tf32(v[4 * q + 1], hi.y, lo.y);
          split_tf32(v[4 * q + 2], hi.z, lo.z);
          split_tf32(v[4 * q + 3], hi.w, lo.w);
          *reinterpret_cast<float4*>(ring_hi + q * S::a_lbo) = hi;
          *reinterpret_cast<float4*>(ring_hi + S::ring_one + q * S::a_lbo) = lo;
        }
        fence_proxy_async();
        fence_before_sync();                             // D1 reads ordered before the hand-off
        nbar_arrive(kPjBarChunk0 + slot, kPjThreads);
      }
      const long long t2 = DENO_CLK();
      if (warp == 1) { DENO_DBG_ADD(22, t2 - t1 - t_slot); DENO_DBG_ADD(23, t_slot); }   // slots 22/23: chunk math + stores, ring waits
      const bool inside_k = inside_cur;
      const int b_k = b_cur, pp_k = pp_cur;
      if (k + 1 < nmine) {                               // GEMM 1 (k) finished long ago: A1 image of the next tile
        pj_store_a1<C, HD, 2>(smem, xrow, wg, hr);
        fence_proxy_async();
        nbar_arrive(kPjBarA1, kPjThreads);
        inside_cur = inside; b_cur = b; pp_cur = pp; gy_cur = gy_next;
        if (k + 2 < nmine) {
          pj_load_tile<C, 2>(p, it.packed(), xrow, wg, hr, inside, valid, b, pp, pc);
          it.next();
          gy_next = valid ? __ldg(p.gy + (long long)b * p.g.npts + pc) : 0.f;
        }
      }
      const long long t3 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(24, t3 - t2);          // slot 24: A1 image of the next tile + prefetch issue
      // the last two chunks' commits: all of GEMM 2 (k) has completed after the second wait
      mbar_wait(&mbar_slot[NCH & 1], slot_phase[NCH & 1]);
      slot_phase[NCH & 1] ^= 1;
      mbar_wait(&mbar_slot[(NCH + 1) & 1], slot_phase[(NCH + 1) & 1]);
      slot_phase[(NCH + 1) & 1] ^= 1;
      fence_after_sync();
      const long long t4 = DENO_CLK();
      if (warp == 1) DENO_DBG_ADD(25, t4 - t3);          // slot 25: wait for the rest of GEMM 2
      float* gh_dst = p.gh + ((long long)b_k * C + wg * (C / 2)) * p.g.nppts + pp_k;
#pragma unroll
      for (int c0 = 0; c0 < C / 2; c0 += 16) {
        float d[16];
        tmem_ld16x3_sum(tmem_d2 + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wg * (C / 2) + c0), (uint32_t)C, d);
        if (inside_k) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) gh_dst[(long long)(c0 + jj) * p.g.nppts] = d[jj];
        }
      }
      fence_before_sync();                               // D2 reads ordered before the next tile's first hand-off
      if (warp == 1) { DENO_DBG_ADD(26, DENO_CLK() - t4); DENO_DBG_ADD(27, 1); }   // slots 26/27: gh epilogue, tiles
    }
    // end of kernel: sum the per-thread accumulators over the 128 lanes of each column half (fixed order)
    float* red = reinterpret_cast<float*>(smem + S::off_red);      // [4 (wq)][2 * HD + 4]
    constexpr int RP = 2 * HD + 4;
#pragma unroll
    for (int i = 0; i < HD / 2; ++i) {
      float s1 = acc_b1[i], s2 = acc_w2[i];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
        s2 += __shfl_xor_sync(0xffffffffu, s2, sh);
      }
      if (lane == 0) {
        red[wq * RP + wg * (HD / 2) + i] = s1;
        red[wq * RP + HD + wg * (HD / 2) + i] = s2;
      }
    }
    {
      float s = acc_b2;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      if (lane == 0 && wg == 0) red[wq * RP + 2 * HD] = s;
    }
    nbar_sync(kPjBarWorkers, kPjWorkers);
    float* out = p.partial + (long long)blockIdx.x * (2 * HD + 1);
    for (int i = tid; i < 2 * HD + 1; i += kPjWorkers)
      out[i] = (red[i] + red[RP + i]) + (red[2 * RP + i] + red[3 * RP + i]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace tc
}  // namespace deno
