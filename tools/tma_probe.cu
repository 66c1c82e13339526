// tma_probe.cu - pins, on hardware, the tensor-map TMA + SWIZZLE_128B + tcgen05 (kind::tf32, K-major) combination the
// kernels of csrc/spectral_tma.cuh rely on.  D[128][N] = A[128][32] * B[N][32]^T with small-integer operands (exact in
// tf32), four K = 8 steps advanced by +32 bytes inside the swizzle atom.  Variants: tile filled by TMA or by threads
// with the documented swizzle formula; LBO field 0 / 1; N = 128 / 192 / 96 / 48; A assembled from four 32-row boxes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  float* out;          // [128][N]
  const float* a_raw;  // [128][32] row-major (manual fill)
  const float* b_raw;  // [N][32]
  int N;
  int use_tma;         // 1: TMA boxes, 0: threads write the swizzled layout
  int lbo;             // LBO field value (16-byte units)
  int a_boxes;         // TMA: A loaded as a_boxes boxes of 128 / a_boxes rows
  int mask_in_place;   // builders mask the tile to tf32 in place before the MMA (generic-proxy writes after TMA)
  int same_acc_twice;  // issue every MMA twice into the same accumulator with half-valued B (tests back-to-back accumulate)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA32,
                                                        const __grid_constant__ CUtensorMap tmB, Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_tma, bar_mma;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) ((float*)smem)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (a.use_tma) {
    if (tid == 0) {
      const uint32_t tx = (uint32_t)((128 + a.N) * 128);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(tx) : "memory");
      if (a.a_boxes == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sA)), "l"(&tmA), "r"(0), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
      } else {
        for (int j = 0; j < 4; ++j)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(smem_u32(sA + j * 32 * 128)), "l"(&tmA32), "r"(0), "r"(j * 32), "r"(smem_u32(&bar_tma)) : "memory");
      }
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sB)), "l"(&tmB), "r"(0), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar_tma)), "r"(0) : "memory");
  } else {
    for (int i = tid; i < 128 * 32; i += 128) {
      const int r = i / 32, k = i % 32;
      *(float*)(sA + r * 128 + ((((k >> 2) ^ (r & 7)) << 4)) + (k & 3) * 4) = a.a_raw[i];
    }
    for (int i = tid; i < a.N * 32; i += 128) {
      const int r = i / 32, k = i % 32;
      *(float*)(sB + r * 128 + ((((k >> 2) ^ (r & 7)) << 4)) + (k & 3) * 4) = a.b_raw[i];
    }
  }
  if (a.mask_in_place) {
    for (int i = tid; i < (128 + a.N) * 32; i += 128) {
      float* p = (float*)smem + i;
      *p = __uint_as_float(__float_as_uint(*p) & 0xFFFFE000u);
    }
  }
  if (a.same_acc_twice) {
    for (int i = tid; i < a.N * 32; i += 128) ((float*)sB)[i] *= 0.5f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t saddr) {
      uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
      d |= (uint64_t)(a.lbo & 0x3FFF) << 16;
      d |= (uint64_t)(1024 >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    uint64_t da = desc(smem_u32(sA)), db = desc(smem_u32(sB));
    for (int s = 0; s < 4; ++s) {
      const uint32_t acc = s > 0 ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      if (a.same_acc_twice)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      da += 2; db += 2;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar_mma)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < a.N; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) a.out[(size_t)tid * a.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeTiledFn fn, CUtensorMap* m, void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes, uint32_t box_rows) {
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {row_bytes};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || f == nullptr) { printf("no cuTensorMapEncodeTiled entry point\n"); return 3; }
  EncodeTiledFn fn = (EncodeTiledFn)f;
  const int smem_bytes = (128 + 256) * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int NMAX = 256, PITCH = 40;     // global rows are 40 floats apart (160 bytes): not the smem pitch
  std::vector<float> hA(128 * PITCH), hB(NMAX * PITCH), rawA(128 * 32), rawB(NMAX * 32);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < PITCH; ++k) hA[r * PITCH + k] = (float)((r * 7 + k * 3) % 11 - 5);
  for (int r = 0; r < NMAX; ++r) for (int k = 0; k < PITCH; ++k) hB[r * PITCH + k] = (float)((r * 5 + k * 2) % 13 - 6);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) rawA[r * 32 + k] = hA[r * PITCH + k];
  for (int r = 0; r < NMAX; ++r) for (int k = 0; k < 32; ++k) rawB[r * 32 + k] = hB[r * PITCH + k];
  float *dA, *dB, *dRawA, *dRawB, *dOut;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dRawA, rawA.size() * 4)); CK(cudaMalloc(&dRawB, rawB.size() * 4));
  CK(cudaMalloc(&dOut, 128 * NMAX * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dRawA, rawA.data(), rawA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dRawB, rawB.data(), rawB.size() * 4, cudaMemcpyHostToDevice));
  struct V { const char* name; int N, use_tma, lbo, a_boxes, mask, twice; };
  const V vs[] = {
      {"manual fill, LBO=1, N=128", 128, 0, 1, 1, 0, 0},
      {"manual fill, LBO=0, N=128", 128, 0, 0, 1, 0, 0},
      {"TMA, LBO=1, N=128", 128, 1, 1, 1, 0, 0},
      {"TMA, LBO=0, N=128", 128, 1, 0, 1, 0, 0},
      {"TMA, A as four 32-row boxes, N=128", 128, 1, 1, 4, 0, 0},
      {"TMA + in-place tf32 mask, N=128", 128, 1, 1, 1, 1, 0},
      {"TMA, N=192", 192, 1, 1, 1, 0, 0},
      {"TMA, N=96", 96, 1, 1, 1, 0, 0},
      {"TMA, N=48", 48, 1, 1, 1, 0, 0},
      {"TMA, N=192, every MMA twice into one accumulator (B halved)", 192, 1, 1, 1, 0, 1},
  };
  for (const V& v : vs) {
    CUtensorMap tmA, tmA32, tmB;
    if (!make_map(fn, &tmA, dA, 32, 128, PITCH * 4, 128)) return 4;
    if (!make_map(fn, &tmA32, dA, 32, 128, PITCH * 4, 32)) return 4;
    if (!make_map(fn, &tmB, dB, 32, NMAX, PITCH * 4, (uint32_t)v.N)) return 4;
    Args a;
    a.out = dOut; a.a_raw = dRawA; a.b_raw = dRawB; a.N = v.N; a.use_tma = v.use_tma; a.lbo = v.lbo; a.a_boxes = v.a_boxes;
    a.mask_in_place = v.mask; a.same_acc_twice = v.twice;
    CK(cudaMemset(dOut, 0, 128 * NMAX * 4));
    probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmA32, tmB, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("== %s : CUDA error %s\n", v.name, cudaGetErrorString(e)); return 5; }
    std::vector<float> out(128 * v.N);
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < v.N; ++n) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)rawA[m * 32 + k] * rawB[n * 32 + k];
        const double d = fabs(ref - out[m * v.N + n]);
        if (d > maxerr) maxerr = d;
        if (d > 1e-3 && bad++ < 4) printf("   mismatch D[%d][%d] = %g, expected %g\n", m, n, out[m * v.N + n], ref);
      }
    printf("== %-62s : %s (max abs err %g, %d bad)\n", v.name, bad ? "WRONG" : "ok", maxerr, bad);
  }
  // does the tensor core truncate or round an fp32 operand to tf32?  A = 1 + 2^-11 + 2^-12 (rounds up under RN), B = 1
  return 0;
}
