// tc_mma_bench3.cu - hardware floor of one tcgen05.mma versus the cost of the loop that issues it.
// tools/tc_mma_bench.cu measured ~74-81 cycles per instruction at every N <= 128; tc_mma_bench2.cu, with more
// arithmetic around the instruction, 149 at every shape: so those numbers are the ISSUING LOOP, not the tensor pipe.
// Here the loop body is NU back-to-back instructions whose descriptors were all computed before the clock starts
// ("pure"), or the product kernels' pattern of 64-bit descriptor adds between instructions ("adds").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_mma_bench3 tools/tc_mma_bench3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
struct Args { int M, N, iters, swz, ks; long long* cycles; };   // ks: distinct K steps walked (operand footprint)
constexpr int NU = 12;

template <bool BF16>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (BF16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t at, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (BF16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: pure (descriptor arrays), 1: adds between instructions, 2: pure with A in tensor memory,
//      3: pure, accumulator alternates between two tiles, 4: the product kernels' 3xTF32 pattern per K step
//      (A_hi x [B_hi|B_lo] -> D0 with N, then A_lo x B_hi -> D1 with N / 2), 5: the same instructions grouped by kind
//      (all first-kind K steps, then all second-kind), 6: pattern 4 with A in tensor memory, 7: pattern 5 with A in tensor memory
template <bool BF16, int MODE>
__global__ void __launch_bounds__(128, 1) bench_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t fmt = BF16 ? 1u : 2u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
  auto desc = [&](uint32_t addr, uint32_t rows) {
    if (a.swz)
      return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t lbo = 16 * rows;
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
  };
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const bool leader = elect_one();
    const uint64_t a0 = desc(smem_u32(smem), 128), b0 = desc(smem_u32(smem + 64 * 1024), a.N);
    const uint64_t ainc = a.swz ? (32 >> 4) : ((2 * 16 * 128) >> 4), binc = a.swz ? (32 >> 4) : ((2 * 16 * a.N) >> 4);
    const uint32_t d0 = tm + 256;
    uint64_t ad[NU], bd[NU];
    uint32_t at[NU];
#pragma unroll
    for (int s = 0; s < NU; ++s) { ad[s] = a0 + (s % a.ks) * ainc; bd[s] = b0 + (s % a.ks) * binc; at[s] = tm + (s % a.ks) * 8; }
    t0 = clock64();
    if (MODE == 1) {
      for (int it = 0; it < a.iters; ++it) {
        uint64_t x = a0, y = b0;
#pragma unroll
        for (int s = 0; s < NU; ++s) {
          if (leader) mma_ss<BF16>(d0, x, y, idesc, (it | s) ? 1u : 0u);
          x += ainc; y += binc;
          if ((s + 1) % a.ks == 0) { x = a0; y = b0; }
        }
      }
    } else if (MODE >= 3) {
      const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(a.N >> 4) << 17);   // N / 2
      const uint32_t d1 = d0 + (uint32_t)a.N;
      uint64_t al[NU];
#pragma unroll
      for (int s = 0; s < NU; ++s) al[s] = ad[s] + (uint64_t)((32 * 1024) >> 4);         // "lo" image 32 KB further
      for (int it = 0; it < a.iters; ++it) {
#pragma unroll
        for (int s = 0; s < NU; ++s) {
          if (leader) {
            if (MODE == 3) mma_ss<BF16>((s & 1) ? d1 : d0, ad[s], bd[s], idesc, 1u);
            if (MODE == 4) { if (s & 1) mma_ss<BF16>(d1, al[s >> 1], bd[s >> 1], idesc2, 1u); else mma_ss<BF16>(d0, ad[s >> 1], bd[s >> 1], idesc, 1u); }
            if (MODE == 5) { if (s >= NU / 2) mma_ss<BF16>(d1, al[s - NU / 2], bd[s - NU / 2], idesc2, 1u); else mma_ss<BF16>(d0, ad[s], bd[s], idesc, 1u); }
            if (MODE == 6) { if (s & 1) mma_ts<BF16>(d1, at[s >> 1] + 64, bd[s >> 1], idesc2, 1u); else mma_ts<BF16>(d0, at[s >> 1], bd[s >> 1], idesc, 1u); }
            if (MODE == 7) { if (s >= NU / 2) mma_ts<BF16>(d1, at[s - NU / 2] + 64, bd[s - NU / 2], idesc2, 1u); else mma_ts<BF16>(d0, at[s], bd[s], idesc, 1u); }
          }
        }
      }
    } else {
      for (int it = 0; it < a.iters; ++it) {
#pragma unroll
        for (int s = 0; s < NU; ++s) {
          if (leader) {
            if (MODE == 2) mma_ts<BF16>(d0, at[s], bd[s], idesc, 1u);
            else mma_ss<BF16>(d0, ad[s], bd[s], idesc, 1u);
          }
        }
      }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    t1 = clock64();
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
  }
  if (tid == 0) {
    const long long t2 = clock64();
    a.cycles[0] = t1 - t0;
    a.cycles[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <bool BF16, int MODE>
void run(const char* kind, const char* mode, int M, int N, int swz, long long* d, int ks = 4) {
  CK(cudaFuncSetAttribute(bench_kernel<BF16, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 16;
  Args a{M, N, iters, swz, ks, d};
  for (int rep = 0; rep < 2; ++rep) {
    bench_kernel<BF16, MODE><<<1, 128, 200 * 1024>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s %s M=%d N=%d swz=%d: %s\n", kind, mode, M, N, swz, cudaGetErrorString(e)); exit(3); }
  }
  long long h[2];
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  printf("%5s %6s ks%-2d %4d %4d %4d | %10lld %12lld %8.1f %8.1f\n", kind, mode, ks, M, N, swz, h[0], h[1], (double)h[0] / (iters * NU),
         (double)h[1] / (iters * NU));
  fflush(stdout);
}

int main() {
  CK(cudaSetDevice(0));
  long long* d;
  CK(cudaMalloc(&d, 16));
  printf("%5s %6s %4s %4s %4s | %10s %12s %8s %8s\n", "kind", "loop", "M", "N", "swz", "issue cyc", "complete cyc", "iss/mma", "cyc/mma");
  for (int M : {128, 64})
    for (int N : {16, 32, 64, 128, 256})
      for (int swz : {0}) {
        run<false, 0>("tf32", "pure", M, N, swz, d);
        run<false, 1>("tf32", "adds", M, N, swz, d);
        if (!swz) run<false, 2>("tf32", "a-tmem", M, N, swz, d);
        if (!swz && M == 128 && N >= 32 && N <= 128) {
          run<false, 3>("tf32", "altD", M, N, swz, d);
          run<false, 4>("tf32", "3x-alt", M, N, swz, d);
          run<false, 5>("tf32", "3x-grp", M, N, swz, d);
          run<false, 6>("tf32", "3x-altT", M, N, swz, d);
          run<false, 7>("tf32", "3x-grpT", M, N, swz, d);
          run<false, 0>("tf32", "pure", M, N, swz, d, 12);
          run<false, 1>("tf32", "adds", M, N, swz, d, 12);
          run<false, 4>("tf32", "3x-alt", M, N, swz, d, 6);
          run<false, 2>("tf32", "a-tmem", M, N, swz, d, 12);
        }
      }
  return 0;
}
