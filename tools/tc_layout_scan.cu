// tc_layout_scan.cu - discovers, on hardware, which (row, k) element of an operand a given shared
// memory word feeds, for a given tcgen05 smem-descriptor (LBO, SBO, major-ness), kind::tf32, M=128.
// Method: A image = zeros except one word = 1.0; B = K-major "identity" (B[n][k] = (n == k)), so
// D[m][n] = A(m, k=n) and the single non-zero of D names the element.  One MMA per scanned word.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_layout_scan tools/tc_layout_scan.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ScanArgs {
  uint64_t a_desc_hi, b_desc_hi;
  uint32_t idesc;
  int scan_a;        // 1: scan the A operand (B is the identity), 0: scan B (A is the identity)
  int nwords;        // words of the scanned image to probe
  int img_bytes;     // size of the scanned image
  int* map;          // [nwords] -> (m << 8) | n of the single nonzero, -1 none, -2 several
};

__global__ void __launch_bounds__(128, 1) scan_kernel(ScanArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int hit_count, hit_code;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* scan_img = (float*)smem;                          // scanned operand image
  float* id_img = (float*)(smem + 64 * 1024);              // identity operand, K-major dense: LBO=128, SBO=256

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // identity operand with R rows (R = 32 if it is B, 128 if it is A), K = 8:
  // elem(r,k) at (r%8)*16 + (r/8)*256 + (k%4)*4 + (k/4)*128 bytes
  const int idR = args.scan_a ? 32 : 128;
  for (int i = tid; i < idR * 8; i += 128) {
    const int r = i / 8, k = i % 8;
    id_img[((r % 8) * 16 + (r / 8) * 256 + (k % 4) * 4 + (k / 4) * 128) / 4] = (r == k) ? 1.0f : 0.0f;
  }
  for (int i = tid; i < args.img_bytes / 4; i += 128) scan_img[i] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  const uint64_t scan_desc = (args.scan_a ? args.a_desc_hi : args.b_desc_hi) | (uint64_t)((smem_u32(scan_img) & 0x3FFFF) >> 4);
  const uint64_t id_desc = (args.scan_a ? args.b_desc_hi : args.a_desc_hi) | (uint64_t)((smem_u32(id_img) & 0x3FFFF) >> 4);
  uint32_t parity = 0;

  for (int w = 0; w < args.nwords; ++w) {
    if (tid == 0) {
      scan_img[w] = 1.0f;
      if (w > 0) scan_img[w - 1] = 0.f;
      hit_count = 0;
      hit_code = -1;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const uint64_t adesc = args.scan_a ? scan_desc : id_desc;
      const uint64_t bdesc = args.scan_a ? id_desc : scan_desc;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(args.idesc), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
      uint32_t done = 0;
      const uint32_t addr = smem_u32(&mbar);
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
      }
      parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) {
      if (__uint_as_float(r[j]) != 0.0f) {
        atomicAdd(&hit_count, 1);
        hit_code = ((warp * 32 + lane) << 8) | j;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) args.map[w] = hit_count == 0 ? -1 : (hit_count == 1 ? hit_code : -2);
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

static uint64_t desc_hi(int lbo_bytes, int sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  struct V { const char* name; int scan_a, mn; int lbo, sbo; int nbytes; };
  V vs[] = {
      {"A K-major  LBO=256 SBO=1024 (control)", 1, 0, 256, 1024, 32768},
      {"A MN-major LBO=256 SBO=1024", 1, 1, 256, 1024, 32768},
      {"A MN-major LBO=1024 SBO=256", 1, 1, 1024, 256, 32768},
      {"B MN-major LBO=256 SBO=1024", 0, 1, 256, 1024, 32768},
  };
  for (const V& v : vs) {
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    const uint64_t dense_k = desc_hi(128, 256);
    if (v.scan_a) { a.a_desc_hi = desc_hi(v.lbo, v.sbo); a.b_desc_hi = dense_k; a.idesc = make_idesc(128, 32, v.mn, 0); }
    else          { a.b_desc_hi = desc_hi(v.lbo, v.sbo); a.a_desc_hi = dense_k; a.idesc = make_idesc(128, 32, 0, v.mn); }
    a.scan_a = v.scan_a;
    a.img_bytes = 48 * 1024;
    a.nwords = v.nbytes / 4;
    CK(cudaMalloc(&a.map, a.nwords * sizeof(int)));
    CK(cudaMemset(a.map, 0xFF, a.nwords * sizeof(int)));
    scan_kernel<<<1, 128, 66 * 1024 + 4096>>>(a);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> map(a.nwords);
    CK(cudaMemcpy(map.data(), a.map, a.nwords * sizeof(int), cudaMemcpyDeviceToHost));
    printf("== %s\n", v.name);
    int shown = 0, hits = 0;
    for (int w = 0; w < a.nwords; ++w) {
      if (map[w] == -1) continue;
      ++hits;
      // when scanning A: D[m][n] = A(m, k=n): row = m, k = n.  When scanning B: D[m][n] = B(n, k=m): row = n, k = m
      const int m = map[w] >> 8, n = map[w] & 0xFF;
      const int row = v.scan_a ? m : n, k = v.scan_a ? n : m;
      if (map[w] == -2) { if (shown < 400) { printf("  byte %5d -> MULTI\n", w * 4); ++shown; } continue; }
      // print a compact subset: first words of every 16-byte chunk and everything below 256 bytes
      if (w * 4 < 160 || (w % 4) == 0) {
        if (shown < 160) { printf("  byte %5d -> row %3d k %d\n", w * 4, row, k); ++shown; }
      }
    }
    printf("  total words that feed the MMA: %d (expected %d)\n", hits, (v.scan_a ? 128 : 32) * 8);
    cudaFree(a.map);
  }
  return 0;
}
