/*
 * deno_fno.h - C ABI of the point-wise ends of an FNO stack (same library: libdeno_spectral.so).
 *
 * Replaces, around the spectral layers of include/deno_spectral.h, the torch op chains of
 *   /root/reference/Models/fno/FNOs.py:61-84   (FNO1d.forward)
 *   /root/reference/Models/fno/FNOs.py:131-152 (FNO2d.forward)
 *   /root/reference/Models/fno/FNOs.py:197-218 (FNO3d.forward)
 *
 *   lift        h = pad(permute(fc0(cat(x, grid))))          lines 136-142 of the 2-D stack
 *   projection  y = fc2(gelu(fc1(permute(crop(h)))))         lines 147-152
 *
 * Conventions are those of deno_spectral.h: plain pointers and sizes, every DEVICE buffer owned by the
 * caller, calls only enqueue kernels on `stream`, non-zero return + deno_last_error() on failure.
 * Weights are the nn.Linear tensors as they are (row-major [out_features, in_features]); x, grid and y are
 * channel-last as the reference passes them, h is channel-first and right-padded as the spectral layers use it.
 */
#ifndef DENO_FNO_H_
#define DENO_FNO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct deno_fno_desc {
  int32_t ndim;        /* 1, 2 or 3 spatial axes */
  int32_t batch;
  int32_t width;       /* C: channels of the spectral stack (fc0 out, fc1 in) */
  int32_t hidden;      /* fc1 out / fc2 in */
  int32_t in_feats;    /* last dimension of x: steps * in_dim */
  int32_t grid_feats;  /* last dimension of grid (the reference passes ndim coordinates) */
  int32_t out_dim;     /* fc2 out */
  int32_t spatial[3];  /* un-padded grid, outermost first; unused entries 0 */
  int32_t padding;     /* right pad of EVERY spatial axis (FNOs.py: self.padding); may be 0 */
} deno_fno_desc;

enum { DENO_FNO_WS_LIFT_BWD = 0, DENO_FNO_WS_PROJ_BWD = 1 };

/* 1 if the fused kernels cover this descriptor (width % 4 == 0 and one of the built widths, in_feats +
 * grid_feats <= 16, out_dim <= 4), else 0 with the reason in deno_last_error().  Host only. */
int deno_fno_supported(const deno_fno_desc* desc);

/* Bytes of scratch the backward calls need (`which`: DENO_FNO_WS_*); 0 on error.  Host only. */
size_t deno_fno_workspace_bytes(const deno_fno_desc* desc, int which);

/*
 *   x     (B, *spatial, in_feats)     grid  (B, *spatial, grid_feats)
 *   w0    (width, in_feats + grid_feats)      b0 (width)
 *   h     (B, width, *(spatial + padding))    written completely (zeros in the padding)
 */
int deno_fno_lift_fwd(const deno_fno_desc* desc, const float* x, const float* grid, const float* w0,
                      const float* b0, float* h, void* stream);

/*
 *   gh    (B, width, *(spatial + padding)) cotangent of h
 *   gx    (B, *spatial, in_feats) or NULL (x is data in one-step training; roll-outs need it)
 *   gw0, gb0  overwritten, either may be NULL only if both are
 */
int deno_fno_lift_bwd(const deno_fno_desc* desc, const float* gh, const float* x, const float* grid,
                      const float* w0, float* gx, float* gw0, float* gb0, void* workspace,
                      size_t workspace_bytes, void* stream);

/*
 *   h     (B, width, *(spatial + padding))
 *   w1 (hidden, width)  b1 (hidden)  w2 (out_dim, hidden)  b2 (out_dim)
 *   y     (B, *spatial, out_dim)
 * The hidden layer uses the erf GELU (F.gelu, FNOs.py:150) and is never written to memory.
 */
int deno_fno_proj_fwd(const deno_fno_desc* desc, const float* h, const float* w1, const float* b1,
                      const float* w2, const float* b2, float* y, void* stream);

/*
 * Gradients of the same call; the hidden layer is recomputed from h.
 *   gy    (B, *spatial, out_dim)
 *   gh    (B, width, *(spatial + padding)), written completely (zeros in the padding)
 *   gw1 (hidden, width)  gb1 (hidden)  gw2 (out_dim, hidden)  gb2 (out_dim): overwritten
 */
int deno_fno_proj_bwd(const deno_fno_desc* desc, const float* gy, const float* h, const float* w1,
                      const float* b1, const float* w2, float* gh, float* gw1, float* gb1, float* gw2,
                      float* gb2, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENO_FNO_H_ */
