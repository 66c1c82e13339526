// tc_mma_bench2.cu - what bounds the ~80 cycles per tcgen05.mma that tools/tc_mma_bench.cu measured for kind::tf32,
// M = 128, no-swizzle K-major operands at every N <= 128?  Same single-issuer loop, timing only (operand values are
// arbitrary), over: operand source (both in shared memory / A in tensor memory), shared-memory layout (no swizzle /
// SWIZZLE_128B), M (128 / 64), N (16 .. 256) and kind (tf32, K = 8 / bf16, K = 16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_mma_bench2 tools/tc_mma_bench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args { int M, N, nmma, nacc, ksteps, a_tmem, swz, bf16; long long* cycles; };

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
  const uint32_t fmt = a.bf16 ? 1u : 2u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
  // no swizzle: core matrices of 8 rows x 16 bytes, SBO = 128 (next 8 rows), LBO = bytes between K chunks of 16 bytes
  // SWIZZLE_128B: rows of 128 bytes, SBO = 1024 (next 8 rows), one instruction's K (32 bytes) = +32 bytes of start address
  auto desc = [&](uint32_t addr, uint32_t rows) {
    if (a.swz)
      return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t lbo = 16 * rows;
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
  };
  const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 64 * 1024);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint64_t a0 = desc(abase, 128), b0 = desc(bbase, a.N);
    // K step of one instruction: 32 bytes.  no swizzle: two 16-byte chunks = 2 * LBO; swizzled: +32 bytes inside the 128-byte row
    const uint64_t ainc = a.swz ? (32 >> 4) : ((2 * 16 * 128) >> 4), binc = a.swz ? (32 >> 4) : ((2 * 16 * a.N) >> 4);
    const uint32_t d0 = tmem + 256;                                   // accumulators: columns 256..511
    const uint32_t akcols = 8;                                        // 32 bytes of K = 8 columns of 32 bits
    t0 = clock64();
    int s = 0, acc_i = 0;
    uint64_t ad = a0, bd = b0;
    uint32_t at = tmem;
    for (int i = 0; i < a.nmma; ++i) {
      const uint32_t d = d0 + (uint32_t)(acc_i * a.N);
      const uint32_t acc = i >= a.nacc ? 1u : 0u;
      if (a.a_tmem) {
        if (a.bf16)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      } else {
        if (a.bf16)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
      ad += ainc; bd += binc; at += akcols;
      if (++s == a.ksteps) { s = 0; ad = a0; bd = b0; at = tmem; }
      if (++acc_i == a.nacc) acc_i = 0;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
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
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d;
  CK(cudaMalloc(&d, 16));
  printf("%5s %4s %4s %5s %4s %4s | %10s %12s %8s\n", "kind", "M", "N", "A", "swz", "nacc", "issue cyc", "complete cyc", "cyc/mma");
  for (int bf16 : {0, 1})
    for (int M : {128, 64})
      for (int N : {16, 32, 64, 128, 256})
        for (int a_tmem : {0, 1})
          for (int swz : {0, 1})
            for (int nacc : {1, 2}) {
              if (nacc * N > 256) continue;
              if (a_tmem && swz && N != 64) continue;          // B layout under a TMEM A: one point is enough
              const int nmma = 96;
              Args a{M, N, nmma, nacc, 4, a_tmem, swz, bf16, d};   // 4 K steps = one 128-byte swizzled row
              for (int rep = 0; rep < 2; ++rep) {
                bench_kernel<<<1, 128, 200 * 1024>>>(a);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("config kind=%d M=%d N=%d A=%d swz=%d: %s\n", bf16, M, N, a_tmem, swz, cudaGetErrorString(e)); return 3; }
              }
              long long h[2];
              CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
              printf("%5s %4d %4d %5s %4d %4d | %10lld %12lld %8.1f\n", bf16 ? "bf16" : "tf32", M, N, a_tmem ? "tmem" : "smem", swz, nacc,
                     h[0], h[1], (double)h[1] / nmma);
              fflush(stdout);
            }
  return 0;
}
