// dp_fused.cuh - the data-parallel exchange fused with the optimiser: ONE kernel per training step does
//     all-reduce(mean) of the flat gradient buffer  +  Adam update  +  broadcast of the new parameters
// over NVLink / NVSwitch multicast memory (replaces ncclAllReduce + adam_flat; reference:
// /root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:272 - DistributedDataParallel's gradient averaging - followed by
// torch.optim.Adam.step()).
//
// Every rank holds its flat gradient buffer g and its flat parameter buffer p in SYMMETRIC memory (same offset on every
// rank, mapped into one multicast address each).  Two-shot over the switch, rank r owns slice r of the flat index range:
//   barrier (all ranks' backward kernels are done: their launches precede this kernel in stream order)
//   for i in slice r:   G  = multimem.ld_reduce.add.v4.f32 [g_mc + i]      (the switch adds the W ranks' copies)
//                       g' = G / W;   Adam on (p[i], m[i], v[i]) with g'    (m, v exist for slice r only: sharded state)
//                       multimem.st.v4.f32 [p_mc + i] = p'                  (the switch writes all W replicas)
//   barrier (every replica of p is complete before any rank's next forward reads it; every rank is done reading g)
// NVLink traffic per GPU and step: 4 n bytes out for the reduction, 4 n bytes in for the broadcast, independent of W -
// against 2 (W-1)/W * 4 n each way for a ring, and two kernel launches + NCCL's own launch latency.
// The cross-rank barrier is the signal-pad protocol of torch's symmetric memory (one 32-bit slot per (block, peer),
// compare-and-swap 0 -> 1 with release by the sender, 1 -> 0 with acquire by the receiver; self-resetting, so the kernel
// can be replayed from a CUDA graph).  Spins are bounded: a rank whose peers never arrive traps instead of hanging.
#pragma once
#include "spectral_common.cuh"

namespace deno {

struct DpFusedParams {
  float* p_local;          // this rank's replica of the parameters (symmetric buffer)
  float* p_mc;             // multicast address of the parameter buffers
  float* g_mc;             // multicast address of the gradient buffers
  float* m;
  float* v;
  long long n;             // floats (multiple of 4)
  int rank, world;
  uint32_t** signal_pads;  // device array: signal pad of every rank
  int write_back_grads;    // debug / tests: also store the averaged gradient into every rank's g (slice owner writes)
  float lr, beta1, beta2, eps, weight_decay;
  const float* step;       // device scalar: t (already incremented for this step)
};

constexpr int kDpThreads = 512;
constexpr int kDpMaxBlocks = 64;            // 64 blocks x 8 ranks = the 512 slots of a 2 KB signal pad (the host clamps to the pad size)
constexpr int kDpInFlight = 4;              // independent multimem reductions per thread: the exchange is bound by NVLink round trips, not bandwidth
constexpr long long kDpSpinLimit = 4000000000LL;   // ~2 s of SM clocks

#ifndef DENO_EMULATE
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t expect, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t expect, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(val) : "memory");
  return old;
}

// Threads 0..world-1 of every block: signal block `blockIdx.x` of peer t, then wait for that peer's signal.
__device__ __forceinline__ void dp_sync_blocks(const DpFusedParams& a) {
  __syncthreads();
  if ((int)threadIdx.x < a.world) {
    const int t = threadIdx.x;
    uint32_t* remote = a.signal_pads[t] + (size_t)blockIdx.x * a.world + a.rank;
    uint32_t* mine = a.signal_pads[a.rank] + (size_t)blockIdx.x * a.world + t;
    const long long t0 = clock64();
    while (cas_release_sys(remote, 0u, 1u) != 0u) {
      if (clock64() - t0 > kDpSpinLimit) __trap();
    }
    while (cas_acquire_sys(mine, 1u, 0u) != 1u) {
      if (clock64() - t0 > kDpSpinLimit) __trap();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float adam_one(float pv, float g, float& m, float& v, const DpFusedParams& a, float step_size,
                                          float inv_sqrt_bc2) {
  const float gv = fmaf(a.weight_decay, pv, g);
  m = fmaf(a.beta1, m, (1.0f - a.beta1) * gv);
  v = fmaf(a.beta2, v, (1.0f - a.beta2) * gv * gv);
  return pv - step_size * (m / fmaf(sqrtf(v), inv_sqrt_bc2, a.eps));
}

__global__ void __launch_bounds__(kDpThreads) dp_allreduce_adam_kernel(DpFusedParams a) {
  const float t = a.step[0];
  const float bc1 = 1.0f - powf(a.beta1, t);
  const float bc2 = 1.0f - powf(a.beta2, t);
  const float step_size = a.lr / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  const float inv_w = 1.0f / (float)a.world;
  const long long n4 = a.n / 4;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank, hi = (lo + per < n4) ? lo + per : n4;

  dp_sync_blocks(a);                                   // every rank's gradients are complete
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += kDpInFlight * stride) {
    float4 gq[kDpInFlight];
#pragma unroll
    for (int u = 0; u < kDpInFlight; ++u) {              // all reductions of this trip in flight before the first is used
      const long long i = i0 + u * stride;
      gq[u] = i < hi ? multimem_ld_reduce_add(a.g_mc + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kDpInFlight; ++u) {
      const long long i = i0 + u * stride;
      if (i >= hi) break;
      float4 g = gq[u];
      g.x *= inv_w; g.y *= inv_w; g.z *= inv_w; g.w *= inv_w;
      float4 pv = reinterpret_cast<const float4*>(a.p_local)[i];
      float4 mv = reinterpret_cast<const float4*>(a.m)[i];
      float4 vv = reinterpret_cast<const float4*>(a.v)[i];
      pv.x = adam_one(pv.x, g.x, mv.x, vv.x, a, step_size, inv_sqrt_bc2);
      pv.y = adam_one(pv.y, g.y, mv.y, vv.y, a, step_size, inv_sqrt_bc2);
      pv.z = adam_one(pv.z, g.z, mv.z, vv.z, a, step_size, inv_sqrt_bc2);
      pv.w = adam_one(pv.w, g.w, mv.w, vv.w, a, step_size, inv_sqrt_bc2);
      reinterpret_cast<float4*>(a.m)[i] = mv;
      reinterpret_cast<float4*>(a.v)[i] = vv;
      multimem_st(a.p_mc + 4 * i, pv);
      if (a.write_back_grads) multimem_st(a.g_mc + 4 * i, g);
    }
  }
  dp_sync_blocks(a);                                   // all replicas written; all ranks done reading gradients
}
#endif

}  // namespace deno
