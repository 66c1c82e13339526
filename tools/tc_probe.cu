// tc_probe.cu - stand-alone probe of the tcgen05 / TMEM programming model on sm_100a.
//
// Pins down, on real hardware, the shared-memory descriptor conventions the tensor-core kernels of
// deno4pytorch_b200/csrc/spectral_tc.cuh rely on (no-swizzle K-major and MN-major operands,
// kind::tf32, M=128, cta_group::1), the TMEM accumulator layout read back with tcgen05.ld, and the
// accuracy of the 3xTF32 split.  Each case builds the smem byte image on the HOST from a layout
// hypothesis, so a wrong hypothesis shows up as a large error instead of a silent assumption.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct ProbeArgs {
  const uint8_t* a_img;   // smem image of A (all K steps, hi then lo when split)
  const uint8_t* b_img;
  int a_bytes, b_bytes;
  uint64_t a_desc_hi;     // descriptor without the start address (LBO/SBO/version/layout)
  uint64_t b_desc_hi;
  uint32_t idesc;
  int nsteps;             // number of MMA instructions
  int a_off[64];          // byte offset of each instruction's A operand inside the A image
  int b_off[64];
  float* d_out;           // [128][N]
  int N;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((args.a_bytes + 1023) / 1024) * 1024;

  for (int i = tid; i < args.a_bytes / 4; i += 128) ((uint32_t*)a_s)[i] = ((const uint32_t*)args.a_img)[i];
  for (int i = tid; i < args.b_bytes / 4; i += 128) ((uint32_t*)b_s)[i] = ((const uint32_t*)args.b_img)[i];
  // generic-proxy writes -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;

  if (tid == 0) {
    for (int s = 0; s < args.nsteps; ++s) {
      const uint64_t adesc = args.a_desc_hi | (uint64_t)((smem_u32(a_s + args.a_off[s]) & 0x3FFFF) >> 4);
      const uint64_t bdesc = args.b_desc_hi | (uint64_t)((smem_u32(b_s + args.b_off[s]) & 0x3FFFF) >> 4);
      const uint32_t accum = s > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(args.idesc), "r"(accum) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(&mbar);
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(addr), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // TMEM -> registers: thread (warp w, lane l) reads TMEM lane 32*w + l, 32 columns at a time
  for (int c0 = 0; c0 < args.N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
    for (int j = 0; j < 32; ++j)
      if (c0 + j < args.N) args.d_out[(warp * 32 + lane) * args.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
  }
}

// ---------------------------------------------------------------------------------------------
static uint64_t desc_hi(int lbo_bytes, int sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0: SWIZZLE_NONE
}

static uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;    // c_format = F32
  d |= 2u << 7;    // a_format = TF32
  d |= 2u << 10;   // b_format = TF32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

static float tf32_trunc(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

// Layout hypotheses.  R = rows along the MN direction (M for A, N for B), K = total K of the image.
// K-major:  elem(r,k) at (r%8)*16 + (r/8)*SBO + (k%4)*4 + (k/4)*LBO
static void fill_kmajor(std::vector<uint8_t>& img, const float* src, int R, int K, int lbo, int sbo, size_t base) {
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      size_t off = base + (size_t)(r % 8) * 16 + (size_t)(r / 8) * sbo + (size_t)(k % 4) * 4 + (size_t)(k / 4) * lbo;
      memcpy(&img[off], &src[r * K + k], 4);
    }
}
// MN-major: elem(r,k) at (r%4)*4 + (r/4)*SBO + (k%8)*16 + (k/8)*LBO
static void fill_mnmajor(std::vector<uint8_t>& img, const float* src, int R, int K, int lbo, int sbo, size_t base) {
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      size_t off = base + (size_t)(r % 4) * 4 + (size_t)(r / 4) * sbo + (size_t)(k % 8) * 16 + (size_t)(k / 8) * lbo;
      memcpy(&img[off], &src[r * K + k], 4);
    }
}

struct Case {
  const char* name;
  int a_mn, b_mn;   // 0 = K-major, 1 = MN-major
  int K;            // total K (multiple of 8)
  int split;        // 1 = plain tf32, 3 = 3xTF32
  int swap;         // swap the LBO / SBO roles (negative control)
  int pad;          // extra bytes added to the MN-direction stride (tests non-dense SBO)
};

int main() {
  const int M = 128, N = 32;
  Case cases[] = {
      {"A K-major,  B K-major,  K=8", 0, 0, 8, 1, 0, 0},
      {"A K-major,  B K-major,  K=8  (LBO/SBO swapped: must FAIL)", 0, 0, 8, 1, 1, 0},
      {"A K-major,  B K-major,  K=32", 0, 0, 32, 1, 0, 0},
      {"A MN-major, B K-major,  K=8", 1, 0, 8, 1, 0, 0},
      {"A MN-major, B K-major,  K=32", 1, 0, 32, 1, 0, 0},
      {"A MN-major, B K-major,  K=32, SBO padded +16B", 1, 0, 32, 1, 0, 16},
      {"A K-major,  B MN-major, K=32", 0, 1, 32, 1, 0, 0},
      {"A MN-major, B MN-major, K=32", 1, 1, 32, 1, 0, 0},
      {"A K-major,  B K-major,  K=32, SBO padded +16B", 0, 0, 32, 1, 0, 16},
      {"A MN-major, B K-major,  K=32, 3xTF32", 1, 0, 32, 3, 0, 0},
      {"A K-major,  B K-major,  K=24, 3xTF32", 0, 0, 24, 3, 0, 0},
  };
  int device = 0;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));

  int nfail_unexpected = 0;
  for (const Case& c : cases) {
    const int K = c.K;
    std::vector<float> A(M * K), B(N * K);   // A[m][k], B[n][k]
    srand(1234 + K);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    std::vector<float> Ahi(A), Alo(A), Bhi(B), Blo(B);
    for (size_t i = 0; i < A.size(); ++i) { Ahi[i] = tf32_trunc(A[i]); Alo[i] = A[i] - Ahi[i]; }
    for (size_t i = 0; i < B.size(); ++i) { Bhi[i] = tf32_trunc(B[i]); Blo[i] = B[i] - Bhi[i]; }

    // strides: core matrices are 128 B; dense packing with the K direction fastest for K-major
    // (LBO = 128, SBO = 128 * K/4) and the MN direction fastest for MN-major (SBO = 128, LBO = 128 * R/4)
    auto strides = [&](int mn_major, int R, int& lbo, int& sbo, int& step_bytes) {
      if (!mn_major) {
        lbo = 128;
        sbo = 128 * (K / 4) + c.pad;
        step_bytes = 2 * lbo;           // one instruction consumes K=8 = two core matrices along K
      } else {
        sbo = 128 + c.pad;
        lbo = sbo * (R / 4);
        step_bytes = lbo;               // one instruction consumes K=8 = one core-matrix row block
      }
    };
    int a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
    strides(c.a_mn, M, a_lbo, a_sbo, a_step);
    strides(c.b_mn, N, b_lbo, b_sbo, b_step);
    const size_t a_one = c.a_mn ? (size_t)a_lbo * (K / 8) : (size_t)a_sbo * (M / 8);
    const size_t b_one = c.b_mn ? (size_t)b_lbo * (K / 8) : (size_t)b_sbo * (N / 8);
    std::vector<uint8_t> a_img(a_one * 2 + 1024, 0), b_img(b_one * 2 + 1024, 0);
    auto fill = [&](std::vector<uint8_t>& img, const float* src, int mn, int R, int lbo, int sbo, size_t base) {
      if (mn) fill_mnmajor(img, src, R, K, lbo, sbo, base);
      else fill_kmajor(img, src, R, K, lbo, sbo, base);
    };
    fill(a_img, c.split == 3 ? Ahi.data() : A.data(), c.a_mn, M, a_lbo, a_sbo, 0);
    fill(a_img, Alo.data(), c.a_mn, M, a_lbo, a_sbo, a_one);
    fill(b_img, c.split == 3 ? Bhi.data() : B.data(), c.b_mn, N, b_lbo, b_sbo, 0);
    fill(b_img, Blo.data(), c.b_mn, N, b_lbo, b_sbo, b_one);

    ProbeArgs args;
    memset(&args, 0, sizeof(args));
    args.a_bytes = (int)a_img.size();
    args.b_bytes = (int)b_img.size();
    args.a_desc_hi = c.swap ? desc_hi(a_sbo, a_lbo) : desc_hi(a_lbo, a_sbo);
    args.b_desc_hi = c.swap ? desc_hi(b_sbo, b_lbo) : desc_hi(b_lbo, b_sbo);
    args.idesc = make_idesc(M, N, c.a_mn, c.b_mn);
    args.N = N;
    int ns = 0;
    const int passes = c.split == 3 ? 3 : 1;
    for (int pass = 0; pass < passes; ++pass) {
      // pass 0: hi*hi, pass 1: lo*hi, pass 2: hi*lo
      const size_t ab = (pass == 1) ? a_one : 0, bb = (pass == 2) ? b_one : 0;
      for (int s = 0; s < K / 8; ++s) {
        args.a_off[ns] = (int)(ab + (size_t)s * a_step);
        args.b_off[ns] = (int)(bb + (size_t)s * b_step);
        ++ns;
      }
    }
    args.nsteps = ns;
    uint8_t *da, *db;
    float* dd;
    CK(cudaMalloc(&da, a_img.size()));
    CK(cudaMalloc(&db, b_img.size()));
    CK(cudaMalloc(&dd, M * N * sizeof(float)));
    CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xFF, M * N * sizeof(float)));
    args.a_img = da; args.b_img = db; args.d_out = dd;
    const size_t smem = ((a_img.size() + 1023) / 1024) * 1024 + b_img.size() + 1024;
    probe_kernel<<<1, 128, smem>>>(args);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dd, M * N * sizeof(float), cudaMemcpyDeviceToHost));
    double num = 0, den = 0, num_tf = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0, ref_tf = 0;
        for (int k = 0; k < K; ++k) {
          ref += (double)A[m * K + k] * (double)B[n * K + k];
          ref_tf += (double)Ahi[m * K + k] * (double)Bhi[n * K + k];
        }
        const double got = D[m * N + n];
        num += (got - ref) * (got - ref);
        num_tf += (got - ref_tf) * (got - ref_tf);
        den += ref * ref;
      }
    const double err = sqrt(num / den), err_tf = sqrt(num_tf / den);
    const double tol = c.split == 3 ? 5e-6 : 2e-3;
    const bool ok = err < tol;
    const bool expected_fail = c.swap != 0;
    printf("%-62s rel-L2 vs fp64 %.3e  vs trunc-tf32 model %.3e  -> %s\n", c.name, err, err_tf,
           ok ? "OK" : (expected_fail ? "fails (expected)" : "FAIL"));
    if (!ok && !expected_fail) ++nfail_unexpected;
    cudaFree(da); cudaFree(db); cudaFree(dd);
  }
  printf("unexpected failures: %d\n", nfail_unexpected);
  return nfail_unexpected ? 1 : 0;
}
