// tma_box_probe.cu - which 2-D tensor-map box loads are legal on this GPU?  One process per variant (an illegal
// instruction kills the context):  tma_box_probe <inner> <rows> <row_bytes> <box_rows> <base_off_bytes> <c0> <c1>
// Loads ONE box [box_rows][32 floats] (SWIZZLE_128B) at coordinates (c0, c1) and checks every element against the
// source (zero where out of bounds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/tma_box_probe tools/tma_box_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) box_kernel(const __grid_constant__ CUtensorMap tm, float* out, int box_rows, int c0, int c1) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 256 * 32; i += 128) ((float*)smem)[i] = -7.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(box_rows * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"(&tm), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = tid; i < box_rows * 32; i += 128) {
    const int r = i / 32, k = i % 32;
    out[i] = *(float*)(smem + r * 128 + ((((k >> 2) ^ (r & 7)) << 4)) + (k & 3) * 4);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 8) { printf("usage\n"); return 1; }
  const long long inner = atoll(argv[1]), rows = atoll(argv[2]), row_bytes = atoll(argv[3]);
  const int box_rows = atoi(argv[4]), base_off = atoi(argv[5]), c0 = atoi(argv[6]), c1 = atoi(argv[7]);
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncodeTiledFn fn = (EncodeTiledFn)f;
  const size_t total = (size_t)base_off + (size_t)rows * row_bytes + 4096;
  std::vector<float> h(total / 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  char* d; float* dout;
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&dout, 256 * 32 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)row_bytes};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d + base_off, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  CK(cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 + 1024));
  box_kernel<<<1, 128, 256 * 128 + 1024>>>(tm, dout, box_rows, c0, c1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("KERNEL ERROR: %s\n", cudaGetErrorString(e)); return 5; }
  std::vector<float> out(box_rows * 32);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int rr = 0; rr < box_rows; ++rr)
    for (int k = 0; k < 32; ++k) {
      const long long gi = c0 + k, gr = (long long)c1 + rr;
      float ref = 0.f;
      if (gi >= 0 && gi < inner && gr >= 0 && gr < rows) ref = h[((size_t)base_off + (size_t)gr * row_bytes) / 4 + gi];
      if (out[rr * 32 + k] != ref && bad++ < 3) printf("  mismatch row %d k %d: %g vs %g\n", rr, k, out[rr * 32 + k], ref);
    }
  printf("ok, %d mismatches\n", bad);
  return 0;
}
