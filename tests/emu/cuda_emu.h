// cuda_emu.h - minimal host-side SIMT emulation used ONLY by the CPU test-suite.
//
// TEST INFRASTRUCTURE.  There is no GPU in the build container, so the index arithmetic,
// shared-memory choreography and launch geometry of the SIMT kernels in
// deno4pytorch_b200/csrc/ are checked here by compiling the very same .cu/.cuh sources
// with g++ (-DDENO_EMULATE) and running every CUDA block as a set of cooperative fibers
// (ucontext): one fiber per CUDA thread, __syncthreads() is a real barrier, dynamic
// shared memory is a NaN-poisoned host buffer with canaries.  The result is loaded only
// by tests/test_emulated_kernels.py.  It is NOT a CPU fallback of the product: the
// Python package never builds, loads or references it, and fails loudly without CUDA.
//
// Not emulated (GPU-only code is compiled out under DENO_EMULATE): tcgen05 / TMEM / TMA /
// mbarrier paths, warp shuffles.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }

namespace emu {
inline uint3_emu g_threadIdx, g_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char* g_dyn_smem = nullptr;

struct BlockRunner {
  static constexpr size_t kStack = 256 * 1024;
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<unsigned char> stacks;
  std::vector<char> done;
  std::vector<uint3_emu> tix;
  int n = 0, cur = -1, arrived = 0, alive = 0;
  unsigned gen = 0;
  const std::function<void()>* body = nullptr;
};
inline BlockRunner g_run;

inline void fiber_entry() {
  (*g_run.body)();
  g_run.done[g_run.cur] = 1;
  g_run.alive--;
  // a thread that exits releases a barrier the remaining threads are waiting on
  if (g_run.alive > 0 && g_run.arrived >= g_run.alive) { g_run.arrived = 0; g_run.gen++; }
  swapcontext(&g_run.ctx[g_run.cur], &g_run.main_ctx);
}

inline void yield_fiber() { swapcontext(&g_run.ctx[g_run.cur], &g_run.main_ctx); }

inline void syncthreads() {
  unsigned my_gen = g_run.gen;
  if (++g_run.arrived >= g_run.alive) {
    g_run.arrived = 0;
    g_run.gen++;
    return;
  }
  while (g_run.gen == my_gen) yield_fiber();
}

inline void run_block(const std::function<void()>& body, dim3 block) {
  BlockRunner& r = g_run;
  r.n = int(block.x * block.y * block.z);
  r.ctx.resize(r.n);
  if (r.stacks.size() < size_t(r.n) * BlockRunner::kStack) r.stacks.resize(size_t(r.n) * BlockRunner::kStack);
  r.done.assign(r.n, 0);
  r.tix.resize(r.n);
  r.arrived = 0;
  r.alive = r.n;
  r.gen = 0;
  r.body = &body;
  for (int t = 0; t < r.n; ++t) {
    r.tix[t] = uint3_emu{unsigned(t) % block.x, (unsigned(t) / block.x) % block.y, unsigned(t) / (block.x * block.y)};
    getcontext(&r.ctx[t]);
    r.ctx[t].uc_stack.ss_sp = r.stacks.data() + size_t(t) * BlockRunner::kStack;
    r.ctx[t].uc_stack.ss_size = BlockRunner::kStack;
    r.ctx[t].uc_link = &r.main_ctx;
    makecontext(&r.ctx[t], (void (*)())fiber_entry, 0);
  }
  long guard = 0;
  while (r.alive > 0) {
    for (int t = 0; t < r.n; ++t) {
      if (r.done[t]) continue;
      r.cur = t;
      g_threadIdx = r.tix[t];
      swapcontext(&r.main_ctx, &r.ctx[t]);
    }
    if (++guard > 50000000L) { fprintf(stderr, "emu: barrier deadlock\n"); abort(); }
  }
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  static const size_t kGuard = 256;
  std::vector<unsigned char> buf(smem_bytes + 2 * kGuard + 64);
  unsigned char* base = buf.data() + kGuard;
  base += (64 - (reinterpret_cast<uintptr_t>(base) & 63)) & 63;
  g_gridDim = grid;
  g_blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::memset(buf.data(), 0xA5, buf.size());
        std::memset(base, 0xFF, smem_bytes);  // NaN poison: reads of unwritten smem surface in results
        g_dyn_smem = base;
        g_blockIdx = uint3_emu{bx, by, bz};
        run_block(body, block);
        for (size_t i = 0; i < kGuard; ++i) {
          if (base[-1 - long(i)] != 0xA5 || base[smem_bytes + i] != 0xA5) {
            fprintf(stderr, "emu: dynamic shared memory overrun (block %u,%u,%u)\n", bx, by, bz);
            abort();
          }
        }
      }
}
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
static inline void __syncthreads() { emu::syncthreads(); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

#define DENO_PDL_SYNC() do { } while (0)
#define DENO_DYN_SMEM(name) unsigned char* name = emu::g_dyn_smem
#define DENO_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
