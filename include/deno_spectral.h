/*
 * deno_spectral.h - C ABI of the B200-native FNO spectral-convolution path.
 *
 * One shared library (libdeno_spectral.so, built by plain nvcc for sm_100a, no torch
 * headers) exports the functions below.  They replace, for the hot path only, the
 * library calls the reference makes from Python:
 *
 *   deno_spectral_fwd  <->  SpectralConv1d/2d/3d.forward
 *                           /root/reference/Models/fno/spectral_layers.py:84-115, 183-208, 299-338
 *                           (conv1x1 + rfftn + corner einsum + zero-pad + irfftn + add + activation)
 *   deno_spectral_bwd  <->  the autograd graph PyTorch records for those lines
 *                           (FftR2CBackward / BmmBackward / FftC2RBackward / ConvolutionBackward /
 *                            GeluBackward), SURVEY.md section 2a
 *
 * Conventions
 *   - every pointer except `desc`, the block-pointer arrays and `deno_spectral_twiddle_fill`'s
 *     buffer is a DEVICE pointer; the caller (PyTorch) owns every buffer, the library never
 *     allocates, frees or synchronises, and keeps no mutable global state;
 *   - activations are contiguous channels-first fp32 `(B, C, *spatial)`;
 *   - spectral weights are the reference's parameters as they sit in memory:
 *     `(Cin, Cout, *modes)` complex64 == `(Cin, Cout, *modes, 2)` fp32, 1 / 2 / 4 blocks in the
 *     reference's order weights1..4 (spectral_layers.py:65-68, 159-169, 272-284);
 *   - `lin_w` is `linear.weight` `(Cout, Cin, 1[,1[,1]])`, `lin_b` is `linear.bias` `(Cout)`;
 *   - a call only enqueues work on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = success; otherwise a DENO_E_* code and deno_last_error() (thread-local text).
 *
 * The reference-side binding (ctypes) is shown in INTEGRATION.md.
 */
#ifndef DENO_SPECTRAL_H_
#define DENO_SPECTRAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DENO_SPECTRAL_ABI_VERSION 1

enum deno_status {
  DENO_OK = 0,
  DENO_E_INVALID = 1,     /* bad descriptor / null pointer / size mismatch            */
  DENO_E_UNSUPPORTED = 2, /* legal for the reference but outside what the kernels do */
  DENO_E_WORKSPACE = 3,   /* workspace_bytes smaller than deno_spectral_workspace_bytes */
  DENO_E_CUDA = 4         /* a CUDA runtime call failed (text has cudaGetErrorString)  */
};

/* Activation table of /root/reference/Models/configs.py:17-19 (None -> SiLU on the Python side). */
enum deno_activation {
  DENO_ACT_IDENTITY = 0, /* used when a caller swaps `.activation` for Identity (Transformers.py:403) */
  DENO_ACT_GELU = 1,     /* erf form, nn.GELU() default */
  DENO_ACT_SILU = 2,
  DENO_ACT_RELU = 3,
  DENO_ACT_TANH = 4,
  DENO_ACT_LEAKYRELU = 5 /* negative_slope 0.01 */
};

/* Shape of one layer call.  `spatial`/`modes` use the first `ndim` entries. */
typedef struct deno_spectral_desc {
  int32_t ndim;       /* 1, 2 or 3                                            */
  int32_t batch;      /* B                                                    */
  int32_t cin, cout;  /* in_dim, out_dim of the layer                         */
  int32_t spatial[3]; /* padded grid sizes as seen by the layer               */
  int32_t modes[3];   /* modes1..3 (retained modes per axis)                  */
  int32_t act;        /* enum deno_activation                                 */
  int32_t flags;      /* reserved, must be 0                                  */
} deno_spectral_desc;

int deno_spectral_abi_version(void);
const char* deno_last_error(void);

/* The DENO_* switches (DENO_SPECTRAL_PATH = auto | simt | tc, and the experiment switches listed in csrc/capi.cu)
 * are read from the environment once, at the first call that needs them.  deno_reload_env() reads them again
 * (tests that flip DENO_SPECTRAL_PATH between cases call it; a product process never needs to). */
int deno_reload_env(void);

/* Number of retained complex modes M (1-D m; 2-D 2*m1*m2; 3-D 4*m1*m2*m3); 0 on a bad descriptor. */
size_t deno_spectral_num_modes(const deno_spectral_desc* desc);

/* Constant DFT tables.  The caller allocates `deno_spectral_twiddle_bytes` on the HOST, lets
 * `deno_spectral_twiddle_fill` write them (computed in fp64, rounded once to fp32), copies the
 * buffer to the device once per (shape, modes) and passes that device pointer as `twiddles`. */
size_t deno_spectral_twiddle_bytes(const deno_spectral_desc* desc);
int deno_spectral_twiddle_fill(const deno_spectral_desc* desc, void* host_buf, size_t bytes);

/* Scratch the caller must provide to fwd / bwd (bytes, device memory, 256-byte aligned). */
size_t deno_spectral_workspace_bytes(const deno_spectral_desc* desc, int backward);

/* Bytes of the packed input spectrum X^ (B, Cin, M) complex64 that fwd can save for bwd. */
size_t deno_spectral_xhat_bytes(const deno_spectral_desc* desc);

/*
 * y = act( irfftn(corner_mix(rfftn(x))) + conv1x1(x) + bias )
 *   x          (B, Cin, *spatial) fp32
 *   w_blocks   HOST array of 1/2/4 DEVICE pointers to the weight blocks
 *   lin_b      may be NULL
 *   y          (B, Cout, *spatial) fp32
 *   z_saved    opaque state for bwd, same shape as y (holds act'(z) at the pre-activation z);
 *              NULL for inference
 *   xhat_saved packed spectrum of x for the weight gradient; NULL for inference
 */
int deno_spectral_fwd(const deno_spectral_desc* desc, const float* x, const void* const* w_blocks,
                      const float* lin_w, const float* lin_b, const void* twiddles, float* y,
                      float* z_saved, void* xhat_saved, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * Gradients of the same call.  Any of gx / gw_blocks / glin_w / glin_b may be NULL to skip it.
 *   gy         (B, Cout, *spatial) fp32 cotangent
 *   gw_blocks  HOST array of DEVICE pointers, same layout as the weights; PyTorch's convention for
 *              complex leaves (grad = dL/dRe + i dL/dIm), overwritten (not accumulated)
 */
int deno_spectral_bwd(const deno_spectral_desc* desc, const float* gy, const float* x,
                      const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                      const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                      float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * deno_spectral_bwd with its parameter-gradient kernels on a second stream.  After gz = gy * act'(z) and its spectrum
 * exist, the backward splits into two independent parts: the input-gradient chain (G^x -> row synthesis -> plane
 * synthesis, which the NEXT layer's backward waits for) and the kernels that only produce gW / glin_w / glin_b.  With a
 * `deno_overlap` the second part is enqueued on `ov->side_stream` (fork and join through the two events, both recorded
 * inside this call: when it returns, `stream` is ordered after everything).  Replaces the serial order in which
 * PyTorch's autograd engine runs the same nodes.  Capturable in a CUDA graph (the side stream
 * joins the capture through ev_fork and leaves it through ev_join).  `ov == NULL` is deno_spectral_bwd.
 * The stream and events are created by deno_overlap_create on the CURRENT device and owned by the caller.
 */
typedef struct deno_overlap {
  void* side_stream; /* cudaStream_t, non-blocking        */
  void* ev_fork;     /* cudaEvent_t, timing disabled      */
  void* ev_join;     /* cudaEvent_t, timing disabled      */
} deno_overlap;

int deno_overlap_create(deno_overlap* out);
int deno_overlap_destroy(deno_overlap* ov);
int deno_spectral_bwd_overlap(const deno_spectral_desc* desc, const float* gy, const float* x,
                              const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                              const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                              float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                              void* stream, const deno_overlap* ov);

/*
 * The 4-D stand-alone layer of the reference (SpectralConv4d, /root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-78, and
 * the `conv_i(x) + w_i(x)` (+ GELU) steps of FNO4d.forward, :118-150): axes (x, y, z, t); BOTH signs of the retained modes
 * on the first two axes, the LOW block [:modes3] only on the third, [:modes4] on the half-spectrum axis; FOUR weight
 * blocks `(Cin, Cout, m1, m2, m3, m4)` complex64 in the reference's order weights1..4 = (+x,+y), (-x,+y), (+x,-y), (-x,-y).
 * Same call contract as the 1/2/3-D entry points above (bypass weight `lin_w` (Cout, Cin), optional bias, activation id;
 * pass a zero `lin_w`, no bias and DENO_ACT_IDENTITY for the bare SpectralConv4d).  Planes are (z, t), batched over
 * (b, x, y); the two leading axes are contracted by two passes of the outer-axis kernels.
 */
typedef struct deno_spectral_desc4 {
  int32_t batch;
  int32_t cin, cout;
  int32_t spatial[4]; /* x, y, z, t                         */
  int32_t modes[4];   /* modes1..4 (modes4 already clamped) */
  int32_t act;        /* enum deno_activation               */
  int32_t flags;      /* reserved, must be 0                */
} deno_spectral_desc4;

size_t deno_spectral4_num_modes(const deno_spectral_desc4* desc);       /* 4 * m1 * m2 * m3 * m4 */
size_t deno_spectral4_twiddle_bytes(const deno_spectral_desc4* desc);
int deno_spectral4_twiddle_fill(const deno_spectral_desc4* desc, void* host_buf, size_t bytes);
size_t deno_spectral4_workspace_bytes(const deno_spectral_desc4* desc, int backward);
int deno_spectral4_fwd(const deno_spectral_desc4* desc, const float* x, const void* const* w_blocks,
                       const float* lin_w, const float* lin_b, const void* twiddles, float* y,
                       float* z_saved, void* xhat_saved, void* workspace, size_t workspace_bytes,
                       void* stream);
/* `ov` may be NULL (one stream); see deno_spectral_bwd_overlap */
int deno_spectral4_bwd(const deno_spectral_desc4* desc, const float* gy, const float* x,
                       const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                       const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                       float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                       void* stream, const deno_overlap* ov);

/*
 * One fused Adam step over flat fp32 buffers (all device pointers, `n` elements each), replacing the per-tensor
 * foreach Adam of the reference demos (`torch.optim.Adam(..., betas=(0.7, 0.9), weight_decay=1e-4)`,
 * /root/reference/Demo/Darcy_2d/run_FNO&UNet.py:209).  Same update rule as torch.optim.Adam (no amsgrad); complex
 * parameters are passed as their real view.  `step` is a DEVICE scalar holding the 1-based step count as float,
 * so the call can be captured in a CUDA graph.
 */
int deno_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, const float* step, void* stream);

/* Number of kernels the last fwd / bwd call on this thread enqueued (bench.py's gpu_launches). */
int deno_spectral_last_launch_count(void);

/*
 * Per-kernel attribution (measurement aid, not used on the product path).  Between
 * deno_profile_begin() and deno_profile_end() every kernel this thread enqueues is bracketed by
 * cudaEventRecord on its stream.  deno_profile_end() synchronises those events and writes up to
 * `capacity` records in launch order; it returns the number of launches seen (may exceed capacity).
 * Not usable during CUDA-graph capture.
 */
typedef struct deno_profile_record {
  char name[32];   /* kernel family, e.g. "plane_synthesis" */
  float ms;        /* device time between the two events */
} deno_profile_record;

int deno_profile_begin(void);
int deno_profile_end(deno_profile_record* records, int capacity);

/* Phase-timing scratch of the warp-specialised kernels (64 x int64 clock64() sums written by CTA 0;
 * slots documented in spectral_tc.cuh).  Copies it to `host`, optionally clears it.  Synchronises the
 * device: measurement aid only.  The probe is compiled only into the -DDENO_PHASE_PROBE variant of the library
 * (tools/phase_probe.py); the product build returns zeros. */
int deno_debug_read(long long* host, int count, int clear);

#ifdef __cplusplus
}
#endif
#endif /* DENO_SPECTRAL_H_ */
