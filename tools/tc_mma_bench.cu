// tc_mma_bench.cu - cycle cost of tcgen05.mma kind::tf32 (M=128, K=8 per instruction, K-major no-swizzle
// operands) as a function of N, of the number of independent accumulator tiles, and of operand reuse.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_mma_bench tools/tc_mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args { int N, nmma, nacc, ksteps, mode; long long* cycles; };
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) bench_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) ((float*)smem)[i] = 0.001f * (i & 127);
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
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t lbo_a = 2048, lbo_b = 16 * a.N;
  auto desc = [](uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
  };
  const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 48 * 1024);
  long long t0 = 0, t1 = 0;
  if (a.mode == 0) {
    if (tid == 0) {
      const uint64_t a0 = desc(abase, lbo_a), b0 = desc(bbase, lbo_b);
      const uint64_t ainc = (2 * lbo_a) >> 4, binc = (2 * lbo_b) >> 4;
      t0 = clock64();
      int s = 0, acc_i = 0;
      uint64_t ad = a0, bd = b0;
      for (int i = 0; i < a.nmma; ++i) {
        const uint32_t d = tmem + (uint32_t)(acc_i * a.N);
        const uint32_t acc = i >= a.nacc ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        ad += ainc; bd += binc;
        if (++s == a.ksteps) { s = 0; ad = a0; bd = b0; }
        if (++acc_i == a.nacc) acc_i = 0;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
      t1 = clock64();
    }
  } else if (warp == 0) {
    // whole warp runs the loop with warp-uniform descriptors; only the MMA itself is predicated on the elected lane
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint64_t a0 = desc(abase, lbo_a), b0 = desc(bbase, lbo_b);
    const uint64_t ainc = (2 * lbo_a) >> 4, binc = (2 * lbo_b) >> 4;
    const bool leader = elect_one();
    t0 = clock64();
    for (int i0 = 0; i0 < a.nmma; i0 += 12) {
      uint64_t ad = a0, bd = b0;
#pragma unroll
      for (int s = 0; s < 12; ++s) {
        const uint32_t d = tm + (uint32_t)(((i0 + s) % 3) * a.N);
        const uint32_t acc = (i0 + s) >= 3 ? 1u : 0u;
        if (leader)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        ad += ainc; bd += binc;
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

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long* d;
  CK(cudaMalloc(&d, 16));
  printf("%6s %6s %6s %6s | %10s %12s %10s\n", "N", "nmma", "nacc", "ksteps", "issue cyc", "complete cyc", "cyc/mma");
  const int Ns[] = {16, 32, 64, 128};
  for (int N : Ns)
    for (int nacc : {1, 2, 3, 4, 6, 8}) {
      if (nacc * N > 512) continue;
      for (int mode : {0, 1}) {
        const int nmma = 96;
        if (mode == 1 && nacc != 3) continue;
        Args a{N, nmma, nacc, 12, mode, d};
        for (int rep = 0; rep < 2; ++rep) {
          bench_kernel<<<1, 128, 100 * 1024>>>(a);
          CK(cudaDeviceSynchronize());
        }
        long long h[2];
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        printf("%6d %6d %6d mode%d | %10lld %12lld %10.1f\n", N, nmma, nacc, mode, h[0], h[1], (double)h[1] / nmma);
      }
    }
  return 0;
}
