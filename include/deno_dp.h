/*
 * deno_dp.h - C ABI of the data-parallel exchange of the FNO training step (same library: libdeno_spectral.so).
 *
 * Replaces, for batch-sharded data parallelism over the GPUs of one NVSwitch box, the pair
 *   DistributedDataParallel's gradient all-reduce   /root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:272
 *   torch.optim.Adam.step()                         /root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:262-266, 75-77
 * by ONE kernel: all-reduce(mean) of the flat gradient buffer through NVLink multicast (multimem.ld_reduce), Adam on this
 * rank's slice, multicast store (multimem.st) of the new parameters into every rank's replica (csrc/dp_fused.cuh).
 *
 * Conventions are those of deno_spectral.h: plain pointers and sizes, every buffer owned by the caller, the call only
 * enqueues a kernel on `stream` (capturable in a CUDA graph), non-zero return + deno_last_error() on failure.
 * The caller provides SYMMETRIC allocations (identical size and offset on every rank, e.g. torch's
 * torch.distributed._symmetric_memory.empty + rendezvous) and their multicast mappings.
 */
#ifndef DENO_DP_H_
#define DENO_DP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 *   p_local     this rank's replica of the flat fp32 parameter buffer (n floats, n % 4 == 0, 16-byte aligned)
 *   p_mc, g_mc  multicast addresses of the parameter / gradient buffers of all `world` ranks
 *   m, v        Adam moments, n floats each, local; only this rank's slice [rank * ceil(n/4/world) * 4, ...) is used
 *   signal_pads DEVICE array of `world` pointers: the 32-bit signal pad of every rank (2 KB = 512 slots, zero
 *               before the first call; the kernel leaves them zero)
 *   step        device scalar: Adam's step count t for this update (float, already incremented)
 *   write_back_grads  non-zero: the averaged gradient is also stored into every rank's gradient buffer (tests, logging)
 * Every rank must enqueue the call once per step, in the same order relative to other uses of the signal pads.
 */
int deno_dp_fused_adam(float* p_local, float* p_mc, float* g_mc, float* m, float* v, size_t n, int rank, int world,
                       void* signal_pads, int write_back_grads, float lr, float beta1, float beta2, float eps,
                       float weight_decay, const float* step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENO_DP_H_ */
