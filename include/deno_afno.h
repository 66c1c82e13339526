/*
 * deno_afno.h - C ABI of the adaptive Fourier layers (same library: libdeno_spectral.so).
 *
 * Replaces the torch op chains of
 *   /root/reference/Models/fno/spectral_layers.py:383-409  (AdaptiveFourier1d.forward)
 *   /root/reference/Models/fno/spectral_layers.py:457-487  (AdaptiveFourier2d.forward)
 *   /root/reference/Models/fno/spectral_layers.py:533-565  (AdaptiveFourier3d.forward)
 *
 *   y = x + irfftn_ortho( softshrink( W2^T act( W1^T rfftn_ortho(x) + b1 ) + b2 ) )
 *
 * with block-diagonal complex weights: the C channels are `num_blocks` blocks of `C / num_blocks`, every bin of the
 * spectrum goes through the same two-layer complex MLP of its block, the activation and the soft-shrink act on real
 * and imaginary parts separately (view_as_real in the reference).  Only hidden_size_factor = 1 runs in the reference
 * (weights2 is contracted over its first block axis, :372-381), so the hidden layer has the block size too.
 *
 * The transforms are the library's DFT kernels over the WHOLE spectrum: every frequency of the leading axes
 * (one-sided table f_k = k, odd sizes included) and the first `modes_last` bins of the half-spectrum axis.
 * `modes_last` is what the reference's `[..., :kept_modes]` slice leaves of that axis:
 *   min( int( (prod(spatial) / 2 + 1) * hard_thresholding_fraction ), spatial[last] / 2 + 1 )   (integer division)
 * bins beyond it are zero in the reference (soft-shrink of zero) and are neither computed nor stored here.
 *
 * Conventions are those of deno_spectral.h: plain pointers and sizes, every DEVICE buffer owned by the caller, calls
 * only enqueue kernels on `stream`, non-zero return + deno_last_error() on failure, no CPU path.
 * Weights are the nn.Parameter tensors as they are: w1, w2 (num_blocks, block, block) complex64 indexed [n][in][out],
 * b1, b2 (num_blocks, block) complex64.
 */
#ifndef DENO_AFNO_H_
#define DENO_AFNO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct deno_afno_desc {
  int32_t ndim;         /* 1, 2 or 3 spatial axes */
  int32_t batch;
  int32_t channels;     /* hidden_size */
  int32_t num_blocks;   /* divides channels */
  int32_t spatial[3];   /* outermost first; unused entries 0 */
  int32_t modes_last;   /* 1 .. spatial[last] / 2 + 1, see above */
  int32_t act;          /* enum deno_activation (deno_spectral.h) between the two block layers */
  float sparsity_threshold; /* lambda of the soft-shrink, >= 0 */
  int32_t flags;        /* reserved, must be 0 */
} deno_afno_desc;

size_t deno_afno_num_modes(const deno_afno_desc* desc);      /* complex bins per (batch entry, channel); 0 on error */
size_t deno_afno_twiddle_bytes(const deno_afno_desc* desc);  /* constant tables (DFT factors + the residual's identity bypass) */
int deno_afno_twiddle_fill(const deno_afno_desc* desc, void* host_buf, size_t bytes);   /* host only */
size_t deno_afno_workspace_bytes(const deno_afno_desc* desc, int backward);

/*
 *   x, y        (B, C, *spatial) fp32
 *   xhat_saved  (B, C, num_modes) complex64 or NULL: the ortho spectrum of x, what the backward needs
 */
int deno_afno_fwd(const deno_afno_desc* desc, const float* x, const void* w1, const void* b1, const void* w2,
                  const void* b2, const void* twiddles, float* y, void* xhat_saved, void* workspace,
                  size_t workspace_bytes, void* stream);

/*
 *   gy (B, C, *spatial) cotangent of y;  gx written completely, or NULL
 *   gw1, gb1, gw2, gb2: overwritten (complex64, PyTorch's conjugate convention), all four or none (NULL)
 * The hidden layer is recomputed from xhat_saved.
 */
int deno_afno_bwd(const deno_afno_desc* desc, const float* gy, const void* xhat_saved, const void* w1,
                  const void* b1, const void* w2, const void* b2, const void* twiddles, float* gx, void* gw1,
                  void* gb1, void* gw2, void* gb2, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENO_AFNO_H_ */
