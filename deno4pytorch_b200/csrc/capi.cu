// capi.cu - the extern "C" boundary of libdeno_spectral.so (see include/deno_spectral.h).
//
// Host-side orchestration only: validates the descriptor, carves the caller's workspace,
// picks launch geometry and enqueues the kernels of spectral_simt.cuh (and, for the shapes it
// covers, the tcgen05 path of spectral_tc.cuh) on the caller's stream.  No allocation, no
// synchronisation, no global mutable state (except the opt-in profiling list).
#include "../../include/deno_spectral.h"
#include "../../include/deno_fno.h"
#include "../../include/deno_dp.h"
#include "../../include/deno_afno.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "spectral_simt.cuh"
#include "fno_pointwise.cuh"
#include "afno_pointwise.cuh"
#include "dp_fused.cuh"
#ifndef DENO_EMULATE
#include "spectral_tc.cuh"
#include "spectral_tma.cuh"
#include "spectral_mix_tc.cuh"
#include "fno_proj_tc.cuh"
#endif

namespace {

using namespace deno;

thread_local std::string g_error;
thread_local int g_launches = 0;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

constexpr size_t kAlign = 256;
constexpr int kWgradTcMaxCtas = 4 * 148;   // partials of bypass_wgrad_tc: 148 persistent CTAs x up to 4 stacked batch entries
constexpr int kMaxSmemBytes = 227 * 1024;
constexpr int kProjTcCtas = 148;           // persistent CTAs of fno_proj_bwd_tc (one partial row each)
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

// ------------------------------------------------------------------------------------------
// Descriptor-derived shape bookkeeping
// ------------------------------------------------------------------------------------------
struct Shape {
  int ndim, B, CI, CO;
  int X, H, W;        // outer (3-D only, else 1), plane rows (1 for 1-D), plane cols
  int mX, mH, mW;     // modes on those axes (unused = 0)
  int KX, KH, MW;     // packed modes per axis (1 when the axis is absent)
  int KX4, KH4, MWp;
  long long S;        // points per channel
  int Q;              // KH * MW: packed modes of one plane
  int M;              // total packed modes
  int act;
  // 4-D stand-alone layer (deno_spectral4_*): two outer axes X1, X2 (X = X1 * X2 planes per batch entry, KX = KX1 * KX2),
  // a ONE-SIDED row axis (KH = mH: the reference keeps [:m3] only) and four weight blocks.  Zero for ndim < 4.
  int X1, X2, mX1, mX2, KX1, KX2, KX14, KX24;
  // adaptive Fourier layers (deno_afno_*): the SIMT transform kernels only - the tensor-core variants have not been
  // measured at full-spectrum shapes (KH = H, MW = W / 2 + 1)
  bool simt_only;
};

inline int num_weight_blocks(const Shape& s) { return s.ndim == 4 ? 4 : (1 << (s.ndim - 1)); }

int parse_desc(const deno_spectral_desc* d, Shape& s) {
  if (d == nullptr) return fail(DENO_E_INVALID, "null descriptor");
  if (d->ndim < 1 || d->ndim > 3) return fail(DENO_E_INVALID, "ndim must be 1, 2 or 3");
  if (d->batch < 1 || d->cin < 1 || d->cout < 1) return fail(DENO_E_INVALID, "batch/cin/cout must be positive");
  if (d->flags != 0) return fail(DENO_E_INVALID, "flags must be 0");
  if (d->act < DENO_ACT_IDENTITY || d->act > DENO_ACT_LEAKYRELU) return fail(DENO_E_INVALID, "unknown activation id");
  for (int a = 0; a < d->ndim; ++a) {
    if (d->spatial[a] < 1 || d->modes[a] < 1) return fail(DENO_E_INVALID, "spatial sizes and modes must be positive");
  }
  const int nd = d->ndim;
  s.ndim = nd;
  s.B = d->batch;
  s.CI = d->cin;
  s.CO = d->cout;
  s.act = d->act;
  s.W = d->spatial[nd - 1];
  s.mW = d->modes[nd - 1];
  s.H = nd >= 2 ? d->spatial[nd - 2] : 1;
  s.mH = nd >= 2 ? d->modes[nd - 2] : 0;
  s.X = nd == 3 ? d->spatial[0] : 1;
  s.mX = nd == 3 ? d->modes[0] : 0;
  if (s.mW > s.W / 2 + 1)
    return fail(DENO_E_INVALID, "modes on the last axis exceed the rfft length N/2+1 (the reference fails in einsum)");
  if (nd == 1 && (s.W % 2) != 0)
    return fail(DENO_E_INVALID, "SpectralConv1d needs an even length (the reference's irfft has no n=, spectral_layers.py:106)");
  if ((nd >= 2 && 2 * s.mH > s.H) || (nd == 3 && 2 * s.mX > s.X))
    return fail(DENO_E_UNSUPPORTED, "2*modes exceeds the grid on a leading axis (overlapping corner blocks)");
  s.KH = nd >= 2 ? 2 * s.mH : 1;
  s.KX = nd == 3 ? 2 * s.mX : 1;
  s.MW = s.mW;
  s.KH4 = round_up(s.KH, 4);
  s.KX4 = round_up(s.KX, 4);
  s.MWp = round_up(s.MW, 4);
  s.S = (long long)s.X * s.H * s.W;
  s.Q = s.KH * s.MW;
  s.M = s.KX * s.Q;
  s.X1 = s.X2 = s.mX1 = s.mX2 = s.KX1 = s.KX2 = s.KX14 = s.KX24 = 0;
  s.simt_only = false;
  return DENO_OK;
}

// SpectralConv4d (/root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-78): axes (x, y, z, t); both signs of the first two,
// the low block [:m3] of the third, [:m4] of the half-spectrum axis.  Planes are (z, t), batched over (b, x, y).
int parse_desc4(const deno_spectral_desc4* d, Shape& s) {
  if (d == nullptr) return fail(DENO_E_INVALID, "null descriptor");
  if (d->batch < 1 || d->cin < 1 || d->cout < 1) return fail(DENO_E_INVALID, "batch/cin/cout must be positive");
  if (d->flags != 0) return fail(DENO_E_INVALID, "flags must be 0");
  if (d->act < DENO_ACT_IDENTITY || d->act > DENO_ACT_LEAKYRELU) return fail(DENO_E_INVALID, "unknown activation id");
  for (int a = 0; a < 4; ++a)
    if (d->spatial[a] < 1 || d->modes[a] < 1) return fail(DENO_E_INVALID, "spatial sizes and modes must be positive");
  s.ndim = 4;
  s.B = d->batch; s.CI = d->cin; s.CO = d->cout; s.act = d->act;
  s.X1 = d->spatial[0]; s.X2 = d->spatial[1]; s.H = d->spatial[2]; s.W = d->spatial[3];
  s.mX1 = d->modes[0]; s.mX2 = d->modes[1]; s.mH = d->modes[2]; s.mW = d->modes[3];
  if (s.mW > s.W / 2 + 1)
    return fail(DENO_E_INVALID, "modes on the last axis exceed the rfft length N/2+1 (the reference fails in einsum)");
  if (s.mH > s.H) return fail(DENO_E_INVALID, "modes3 exceeds the third axis (the reference fails in einsum)");
  if (2 * s.mX1 > s.X1 || 2 * s.mX2 > s.X2)
    return fail(DENO_E_UNSUPPORTED, "2*modes exceeds the grid on a leading axis (overlapping corner blocks)");
  s.X = s.X1 * s.X2; s.mX = 0;
  s.KX1 = 2 * s.mX1; s.KX2 = 2 * s.mX2; s.KX = s.KX1 * s.KX2;
  s.KX14 = round_up(s.KX1, 4); s.KX24 = round_up(s.KX2, 4); s.KX4 = round_up(s.KX, 4);
  s.KH = s.mH; s.MW = s.mW;
  s.KH4 = round_up(s.KH, 4); s.MWp = round_up(s.MW, 4);
  s.S = (long long)s.X * s.H * s.W;
  s.Q = s.KH * s.MW;
  s.M = s.KX * s.Q;
  s.simt_only = false;
  return DENO_OK;
}

PlaneGeom make_geom(const Shape& s, int channels) {
  PlaneGeom g;
  g.H = s.H;
  g.W = s.W;
  g.KH = s.KH;
  g.MW = s.MW;
  g.C = channels;
  g.NBO = s.B;
  g.NBI = s.X;
  g.sbo = (long long)channels * s.S;
  g.sbi = (long long)s.H * s.W;
  g.sc = s.S;
  return g;
}

ModeDecode make_decode(const Shape& s) {
  ModeDecode md;
  md.ndim = s.ndim;
  if (s.ndim == 1) {
    md.m1 = s.mW; md.m2 = 1; md.m3 = 1;
    md.Mblk = s.mW;
  } else if (s.ndim == 2) {
    md.m1 = s.mH; md.m2 = s.mW; md.m3 = 1;
    md.Mblk = s.mH * s.mW;
  } else if (s.ndim == 3) {
    md.m1 = s.mX; md.m2 = s.mH; md.m3 = s.mW;
    md.Mblk = s.mX * s.mH * s.mW;
  } else {
    // 4-D: packed mode = ((k1 * KX2 + k2) * m3 + kz) * m4 + l; the blocks depend on (k1, k2) only and a block is
    // [m1][m2][m3 * m4] - exactly the 3-D decode with the (z, t) modes merged into its innermost index
    md.ndim = 3;
    md.m1 = s.mX1; md.m2 = s.mX2; md.m3 = s.mH * s.mW;
    md.Mblk = s.mX1 * s.mX2 * s.mH * s.mW;
  }
  md.M = s.M;
  return md;
}

// ------------------------------------------------------------------------------------------
// Twiddle tables (host, fp64 -> fp32)
// ------------------------------------------------------------------------------------------
struct TwLayout {
  size_t last, mid, outer, outer2, cl, synth_ae, an_e1, an_a2, at_e1, at_a2, total;  // float offsets
  int nchunks, KE;                               // tensor-core synthesis images: [chunk][hi|lo][KE*128]
  int N1, Wp, Hp, KS0, KRp;                      // tensor-core analysis images: E1 [2*N1 rows: hi, lo][Wp], A2 [slab][hi|lo][KRp*Hp]
  int an_nsub, an_HS;                            // row slabs of the tensor-core analysis (0: shape not covered); Hp = an_HS
  size_t rs_a;                                   // row_synthesis_tc A images [rs_ntm][hi|lo][rs_KP * 128] (float offset)
  int rs_KP, rs_ntm;                             // 2 * round_up(KH, 4) columns; 128-row tiles (0: no row axis / too tall)
};

// Shared-memory bytes of plane_analysis_tc_kernel for slabs of HS rows (mirrors tc::analysis_tc_smem, which the launcher
// checks against; kept here because the twiddle layout is also needed by the CPU emulation build).
inline long long an_tc_smem_total(int HS, int W, int KH, int N1, int Wp, int KRp, int nsub) {
  const long long e1 = 2LL * N1 * Wp * 4, a2_one = (long long)KRp * HS * 4;
  const long long a_one = (long long)(Wp / 4) * (16 * HS + 16), b2 = (long long)(HS / 4) * (16 * 2 * N1 + 16);
  const long long xch = round_up(KH * (N1 + 1) * 4, 16), stage = (long long)round_up(HS * W, 4) * 4;
  return e1 + nsub * 2 * a2_one + 2 * a_one + b2 + xch + stage + 4096;
}

// Shape-only geometry of the TMA analysis kernel (independent of the planes-per-tile choice).
// Row class j = row mod g starts j * W floats into the plane; a TMA box must start on a 16-byte boundary in global
// memory (an inner coordinate that is not a multiple of 4 floats is an illegal instruction on hardware:
// profiles/r2g_tma_box_probe.txt), so class j is described from the 16-byte boundary below its first row and the
// `shift[j]` leading floats of every landed row (the tail of the previous row) meet zero columns of that class's
// twiddle image.
struct AtGeom { bool ok; int g, HG, HGp, N1, KHp, nchunks, K2p; int shift[4]; };
AtGeom at_geom(const Shape& s) {
  AtGeom a;
  memset(&a, 0, sizeof(a));
  if (s.ndim < 2) return a;
  const int r = s.W % 4;
  a.g = r == 0 ? 1 : (r == 2 ? 2 : 4);
  if (s.H % a.g != 0) return a;
  a.HG = s.H / a.g;
  a.HGp = round_up(a.HG, 16);
  a.N1 = round_up(2 * s.MW, 8);
  a.KHp = round_up(s.KH, 8);
  int max_shift = 0;
  for (int j = 0; j < a.g; ++j) {
    a.shift[j] = (int)(((long long)j * s.W) % 4);
    if (a.shift[j] > max_shift) max_shift = a.shift[j];
  }
  a.nchunks = (s.W + max_shift + 31) / 32;
  a.K2p = round_up(a.g * a.HGp, 32);
  if (a.N1 > 32 || 4 * a.KHp > 128 || a.HG > 256 || a.g * a.HGp > 256) return a;
  a.ok = true;
  return a;
}

TwLayout tw_layout(const Shape& s) {
  TwLayout t;
  size_t off = 0;
  t.last = off;  off += (size_t)s.W * s.MWp * 2;
  t.mid = off;   off += (size_t)s.H * s.KH4 * 2;
  t.outer = off; off += (s.ndim == 3) ? (size_t)s.X * s.KX4 * 2 : (s.ndim == 4 ? (size_t)s.X1 * s.KX14 * 2 : 0);
  t.outer2 = off; off += (s.ndim == 4) ? (size_t)s.X2 * s.KX24 * 2 : 0;
  t.cl = off;    off += (size_t)s.MWp;
  t.nchunks = (s.W + 127) / 128;
  t.KE = 2 * s.MWp;
  t.synth_ae = off; off += (size_t)t.nchunks * 2 * t.KE * 128;
  t.N1 = round_up(2 * s.MWp, 32);
  t.Wp = round_up(s.W, 8);
  t.Hp = round_up(s.H, 8);
  t.KS0 = s.KH;                                  // sin rows of the stage-2 operand follow the cos rows
  t.KRp = round_up(t.KS0 + s.KH, 8);
  // shapes the tensor-core analysis kernel can take: planes of up to 128 rows as they are, taller ones (Airfoil 229,
  // Pipe 137) as the fewest row slabs of a multiple of 8 rows that fit shared memory
  t.an_nsub = 0; t.an_HS = 0;
  if (s.ndim >= 2 && t.KS0 + s.KH <= 128 && 6 * t.N1 <= 512) {
    for (int n = (s.H + 127) / 128; n <= 4 && t.an_nsub == 0; ++n) {
      const int HS = round_up((s.H + n - 1) / n, 8);
      if (HS > 128 || (n - 1) * HS >= s.H) continue;
      if (an_tc_smem_total(HS, s.W, s.KH, t.N1, t.Wp, t.KRp, n) <= kMaxSmemBytes - 1024) { t.an_nsub = n; t.an_HS = HS; }
    }
  }
  const bool an = t.an_nsub > 0;
  if (an) t.Hp = t.an_HS;
  t.an_e1 = off; off += an ? (size_t)2 * t.N1 * t.Wp : 0;
  t.an_a2 = off; off += an ? (size_t)t.an_nsub * 2 * t.KRp * t.Hp : 0;
  // TMA analysis kernel (spectral_tma.cuh): E image nchunks x [2 N1 rows][32], A2 image (K2p / 32) x [4 KHp rows][32],
  // both already in their SWIZZLE_128B shared-memory layout
  const AtGeom ag = at_geom(s);
  off = (off + 63) / 64 * 64;                      // 256-byte aligned: the images are bulk-copied
  t.at_e1 = off; off += ag.ok ? (size_t)ag.g * ag.nchunks * 2 * ag.N1 * 32 : 0;
  t.at_a2 = off; off += ag.ok ? (size_t)(ag.K2p / 32) * 4 * ag.KHp * 32 : 0;
  // row_synthesis_tc: cos / sin of the row axis as tcgen05 A operand images (K-major, 128 rows per tile)
  t.rs_KP = 2 * s.KH4;
  t.rs_ntm = (s.ndim >= 2 && s.H <= 256) ? (s.H + 127) / 128 : 0;
  off = (off + 63) / 64 * 64;
  t.rs_a = off; off += (size_t)t.rs_ntm * 2 * t.rs_KP * 128;
  t.total = off;
  return t;
}

float tf32_mask(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

// Last-axis twiddles as tcgen05 A-operand images (K-major, SBO = 128, LBO = 2048), split hi/lo from fp64:
// column k = 2l holds cos(2 pi l y / W), k = 2l+1 holds sin(2 pi l y / W); zero outside the grid / modes.
void fill_synth_ae(float* dst, const Shape& s, const TwLayout& t) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int c = 0; c < t.nchunks; ++c) {
    float* hi = dst + (size_t)c * 2 * t.KE * 128;
    float* lo = hi + (size_t)t.KE * 128;
    for (int yl = 0; yl < 128; ++yl) {
      for (int k = 0; k < t.KE; ++k) {
        const int y = c * 128 + yl, l = k / 2;
        double v = 0.0;
        if (y < s.W && l < s.MW) {
          const long long r = ((long long)l * y) % s.W;
          const double ang = two_pi * (double)r / (double)s.W;
          v = (k & 1) ? sin(ang) : cos(ang);
        }
        const float h = tf32_mask((float)v);
        const int off = kmajor_off(yl, k, 128) / 4;
        hi[off] = h;
        lo[off] = (float)(v - (double)h);
      }
    }
  }
}

void fill_axis(float* dst, int n, int nfreq, int pitch, int m, bool half) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int x = 0; x < n; ++x) {
    for (int k = 0; k < pitch; ++k) {
      float c = 0.f, sn = 0.f;
      if (k < nfreq) {
        long long f = half ? k : (k < m ? k : (long long)n - 2 * m + k);
        if (n == 1) f = 0;
        long long r = (f * x) % n;
        double ang = two_pi * (double)r / (double)n;
        c = (float)cos(ang);
        sn = (float)sin(ang);
      }
      dst[((size_t)x * pitch + k) * 2 + 0] = c;
      dst[((size_t)x * pitch + k) * 2 + 1] = sn;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Launch-geometry heuristics (pure functions of the shape so that workspace sizing agrees)
// ------------------------------------------------------------------------------------------
struct AnalysisCfg { int G, RT, RTP, YC, RS, CS, threads; size_t smem; };

int analysis_cfg(const PlaneGeom& g, AnalysisCfg& c) {
  const int MWp = round_up(g.MW, 4), NLG = MWp / 4, KQ = round_up(g.KH, 4) / 4;
  c.G = (g.H >= 32) ? 1 : (g.H == 1 ? 8 : (32 + g.H - 1) / g.H);
  const int rows = c.G * g.H;
  int rtcap = (512 / NLG) & ~31;
  if (rtcap < 32) rtcap = 32;
  const int ntiles = (rows + rtcap - 1) / rtcap;
  c.RT = (rows + ntiles - 1) / ntiles;
  c.RTP = c.RT < 32 ? c.RT : round_up(c.RT, 32);
  // few rows per tile (1-D layers: 8 single-row planes per CTA): split the columns of stage 1 over up to 8 thread
  // slices, otherwise a CTA is one busy warp walking 1 000 columns (178 us for the Burgers layer)
  c.CS = 1;
  while (c.CS < 8 && c.RTP * NLG * c.CS * 2 <= 256) c.CS *= 2;
  c.threads = round_up(c.RTP * NLG * c.CS, 32);
  if (c.threads < 64) c.threads = 64;
  if (c.threads > 1024) return fail(DENO_E_UNSUPPORTED, "too many retained modes on the last axis for plane_analysis");
  const int nchunks = (g.W + 127) / 128;
  c.YC = (g.W + nchunks - 1) / nchunks;
  const int ntask2 = c.G * KQ * g.MW;
  c.RS = c.threads / ntask2;
  if (c.RS < 1) c.RS = 1;
  if (c.RS > 8) c.RS = 8;
  AnalysisParams p;
  p.g = g; p.G = c.G; p.RT = c.RT; p.YC = c.YC; p.RS = c.RS; p.CS = c.CS;
  c.smem = (size_t)analysis_smem(p).total * sizeof(float);
  if (c.smem > (size_t)kMaxSmemBytes) return fail(DENO_E_UNSUPPORTED, "plane_analysis tile does not fit shared memory");
  return DENO_OK;
}

struct SynthCfg { int R, PTS, tiles, threads; size_t smem; };

int synth_cfg(const PlaneGeom& g, int CO, SynthCfg& c) {
  SynthParams p;
  p.g = g;
  p.CO = CO;
  if (g.H == 1) {
    const int nt = (g.W + 255) / 256;
    c.R = 1;
    c.PTS = round_up((g.W + nt - 1) / nt, 4);
    c.tiles = (g.W + c.PTS - 1) / c.PTS;
    p.R = c.R; p.PTS = c.PTS;
    c.smem = (size_t)synth_smem(p).total * sizeof(float);
  } else {
    // pick the rows-per-tile that wastes the fewest CTA slots in the last wave (148 SMs x resident CTAs)
    int best = 0;
    double best_eff = -1.0;
    for (int R = kSynthMaxRows; R >= 1; --R) {
      if (R > g.H) continue;
      p.R = R; p.PTS = R * g.W;
      const size_t bytes = (size_t)synth_smem(p).total * sizeof(float);
      if (bytes > (size_t)kMaxSmemBytes) continue;
      int per_sm = (int)((228 * 1024) / (bytes + 1024));
      if (per_sm > 8) per_sm = 8;          // 256 threads per CTA, 2048 per SM
      if (per_sm < 1) per_sm = 1;
      const long long ctas = (long long)g.NBO * g.NBI * ((g.H + R - 1) / R);
      const long long slots = 148LL * per_sm;
      const long long waves = (ctas + slots - 1) / slots;
      double eff = (double)ctas / (double)(waves * slots);
      if (per_sm < 2) eff *= 0.7;          // a lone CTA per SM cannot overlap its phases
      if (eff > best_eff + 1e-9) { best_eff = eff; best = R; }
    }
    if (best == 0) best = 1;
    c.R = best;
    c.PTS = best * g.W;
    c.tiles = (g.H + best - 1) / best;
    p.R = c.R; p.PTS = c.PTS;
    c.smem = (size_t)synth_smem(p).total * sizeof(float);
  }
  if (c.smem > (size_t)kMaxSmemBytes) return fail(DENO_E_UNSUPPORTED, "plane_synthesis tile does not fit shared memory (too many channels)");
  c.threads = 256;
  return DENO_OK;
}

struct WgradCfg { int PTS, tiles, ncta, threads; size_t smem; };

int wgrad_cfg(const Shape& s, WgradCfg& c) {
  const int rows = s.CO + round_up(s.CI, 8) + 1;
  int pts = ((96 * 1024 / 4) / rows - 4) / 32 * 32;
  if (pts > 512) pts = 512;
  if (pts < 32) return fail(DENO_E_UNSUPPORTED, "too many channels for bypass_wgrad");
  const int plane = s.H * s.W;
  if (pts > round_up(plane, 32)) pts = round_up(plane, 32);
  c.PTS = pts;
  c.tiles = (plane + pts - 1) / pts;
  c.ncta = s.B * s.X * c.tiles;
  c.threads = 256;
  c.smem = (size_t)rows * (pts + 4) * sizeof(float);
  return DENO_OK;
}

// ------------------------------------------------------------------------------------------
// Workspace carving
// ------------------------------------------------------------------------------------------
struct FwdWs { size_t xhat, yhat, t2, v, t3, v3, urow, total; };
struct BwdWs { size_t gz, ghat, gxhat, t2, v, t3, v3, partial, urow, total; };

// pre-split row-synthesis images for the tensor-core synthesis kernel: [B*X][H][hi|lo][KE * C] floats
size_t urow_bytes(const Shape& s, int channels) {
  if (s.ndim < 2) return 0;
  return align_up((size_t)s.B * s.X * s.H * 2 * (2 * (size_t)s.MWp) * channels * 4);
}

FwdWs fwd_ws(const Shape& s) {
  FwdWs w;
  size_t off = 0;
  w.xhat = off; off += align_up((size_t)s.B * s.CI * s.M * 8);
  w.yhat = off; off += align_up((size_t)s.B * s.CO * s.M * 8);
  w.t2 = off;   off += (s.ndim >= 3) ? align_up((size_t)s.B * s.X * s.CI * s.Q * 8) : 0;
  w.v = off;    off += (s.ndim >= 3) ? align_up((size_t)s.B * s.X * s.CO * s.Q * 8) : 0;
  w.t3 = off;   off += (s.ndim == 4) ? align_up((size_t)s.B * s.X1 * s.CI * s.KX2 * s.Q * 8) : 0;   // after the first outer pass
  w.v3 = off;   off += (s.ndim == 4) ? align_up((size_t)s.B * s.X1 * s.CO * s.KX2 * s.Q * 8) : 0;
  w.urow = off; off += urow_bytes(s, s.CO);
  w.total = off;
  return w;
}

int bwd_ws(const Shape& s, BwdWs& w) {
  WgradCfg wc;
  int rc = wgrad_cfg(s, wc);
  if (rc != DENO_OK) return rc;
  size_t off = 0;
  w.gz = off;      off += align_up((size_t)s.B * s.CO * s.S * 4);
  w.ghat = off;    off += align_up((size_t)s.B * s.CO * s.M * 8);
  w.gxhat = off;   off += align_up((size_t)s.B * s.CI * s.M * 8);
  w.t2 = off;      off += (s.ndim >= 3) ? align_up((size_t)s.B * s.X * s.CO * s.Q * 8) : 0;
  w.v = off;       off += (s.ndim >= 3) ? align_up((size_t)s.B * s.X * s.CI * s.Q * 8) : 0;
  w.t3 = off;      off += (s.ndim == 4) ? align_up((size_t)s.B * s.X1 * s.CO * s.KX2 * s.Q * 8) : 0;
  w.v3 = off;      off += (s.ndim == 4) ? align_up((size_t)s.B * s.X1 * s.CI * s.KX2 * s.Q * 8) : 0;
  const size_t nparts = (size_t)(wc.ncta > kWgradTcMaxCtas ? wc.ncta : kWgradTcMaxCtas);   // either kernel's partial count
  w.partial = off; off += align_up(nparts * ((size_t)s.CO * s.CI + s.CO) * 4);
  w.urow = off;    off += urow_bytes(s, s.CI);
  w.total = off;
  return DENO_OK;
}

// ------------------------------------------------------------------------------------------
// Launch helpers
// ------------------------------------------------------------------------------------------
// Optional per-launch event bracketing (deno_profile_begin / deno_profile_end).
struct ProfEntry {
  const char* name;
#ifndef DENO_EMULATE
  cudaEvent_t start, stop;
#endif
};
// Process-wide (autograd runs backward on its own thread), guarded by a mutex; off by default.
bool g_profiling = false;
std::vector<ProfEntry> g_prof;
std::mutex g_prof_mutex;

void prof_before(const char* what, cudaStream_t st) {
  if (!g_profiling) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  ProfEntry e;
  e.name = what;
#ifndef DENO_EMULATE
  cudaEventCreate(&e.start);
  cudaEventCreate(&e.stop);
  cudaEventRecord(e.start, st);
#else
  (void)st;
#endif
  g_prof.push_back(e);
}

void prof_after(cudaStream_t st) {
  if (!g_profiling) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (g_prof.empty()) return;
#ifndef DENO_EMULATE
  cudaEventRecord(g_prof.back().stop, st);
#else
  (void)st;
#endif
}

int check_launch(const char* what, cudaStream_t st) {
  ++g_launches;
  prof_after(st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DENO_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return DENO_OK;
}

template <typename K>
int allow_smem(K kernel, size_t bytes, const char* what) {
#ifndef DENO_EMULATE
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(DENO_E_CUDA, std::string(what) + " smem opt-in: " + cudaGetErrorString(e));
  }
#else
  (void)kernel; (void)bytes; (void)what;
#endif
  return DENO_OK;
}

#ifndef DENO_EMULATE
// ------------------------------------------------------------------------------------------
// Tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// fp32 matrix [outer][inner] with `row_bytes` between rows, boxes of [box_outer][32 floats], SWIZZLE_128B, zero OOB fill
bool make_tmap_2d(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long outer,
                  unsigned long long row_bytes, unsigned box_outer) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_bytes & 15) != 0 || box_outer == 0 || box_outer > 256) return false;
  const cuuint64_t gdim[2] = {inner, outer};
  const cuuint64_t gstr[1] = {row_bytes};
  const cuuint32_t box[2] = {32, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#endif

// Analysis operand images (see plane_analysis_tc_kernel): E1[n][y] (n = 2l: cos, 2l+1: -sin of 2 pi l y / W)
// as ONE B operand image with 2 N1 rows (tf32 hi parts, then lo parts), and A2[r][x] (rows r < KH: cos, rows
// KS0 + k: sin of 2 pi f_k x / H) as hi / lo A operand images.
void fill_analysis_images(float* base, const Shape& s, const TwLayout& t) {
  const double two_pi = 6.283185307179586476925286766559;
  if (t.an_nsub == 0) return;
  float* e1 = base + t.an_e1;                    // one B image: rows n (hi) and N1 + n (lo)
  for (int n = 0; n < t.N1; ++n)
    for (int y = 0; y < t.Wp; ++y) {
      const int l = n / 2;
      double v = 0.0;
      if (l < s.MW && y < s.W) {
        const double ang = two_pi * (double)(((long long)l * y) % s.W) / (double)s.W;
        v = (n & 1) ? -sin(ang) : cos(ang);
      }
      const float h = tf32_mask((float)v);
      e1[kmajor_off(n, y, 2 * t.N1) / 4] = h;
      e1[kmajor_off(t.N1 + n, y, 2 * t.N1) / 4] = (float)(v - (double)h);
    }
  for (int sub = 0; sub < t.an_nsub; ++sub) {      // one [hi | lo] image pair per row slab: local row xl is plane row sub * HS + xl
    float* a2hi = base + t.an_a2 + (size_t)sub * 2 * t.KRp * t.Hp;
    float* a2lo = a2hi + (size_t)t.KRp * t.Hp;
    for (int r = 0; r < t.KRp; ++r)
      for (int xl = 0; xl < t.Hp; ++xl) {
        const bool is_cos = r < s.KH, is_sin = r >= t.KS0 && r < t.KS0 + s.KH;
        const int k = is_cos ? r : r - t.KS0;
        const long long x = (long long)sub * t.an_HS + xl;
        double v = 0.0;
        if ((is_cos || is_sin) && x < s.H) {
          const long long f = (k < s.mH) ? k : (long long)s.H - 2 * s.mH + k;
          const double ang = two_pi * (double)((f * x) % s.H) / (double)s.H;
          v = is_cos ? cos(ang) : sin(ang);
        }
        const float h = tf32_mask((float)v);
        const int off = kmajor_off(r, xl, t.KRp) / 4;
        a2hi[off] = h;
        a2lo[off] = (float)(v - (double)h);
      }
  }
}

// row_synthesis_tc A operand: A[x, k] = cos(2 pi f_k x / H), A[x, KP/2 + k] = sin(...), tf32 hi / lo from fp64, zero
// outside the grid rows and the retained modes.
void fill_row_synth_images(float* base, const Shape& s, const TwLayout& t) {
  const double two_pi = 6.283185307179586476925286766559;
  const int KH2 = t.rs_KP / 2;
  for (int tm = 0; tm < t.rs_ntm; ++tm) {
    float* hi = base + t.rs_a + (size_t)tm * 2 * t.rs_KP * 128;
    float* lo = hi + (size_t)t.rs_KP * 128;
    for (int r = 0; r < 128; ++r)
      for (int kp = 0; kp < t.rs_KP; ++kp) {
        const int k = kp < KH2 ? kp : kp - KH2;
        const long long x = (long long)tm * 128 + r;
        double v = 0.0;
        if (k < s.KH && x < s.H) {
          const long long f = (k < s.mH) ? k : (long long)s.H - 2 * s.mH + k;
          const double ang = two_pi * (double)((f * x) % s.H) / (double)s.H;
          v = kp < KH2 ? cos(ang) : sin(ang);
        }
        const float h = tf32_mask((float)v);
        const int off = kmajor_off(r, kp, 128) / 4;
        hi[off] = h;
        lo[off] = (float)(v - (double)h);
      }
  }
}

// byte offset of element (row, k) of a K-major SWIZZLE_128B image made of [rows][32 fp32] chunks
inline size_t sw128_off(int row, int k, int rows) {
  return (size_t)(k >> 5) * rows * 128 + (size_t)row * 128 + (size_t)((((k & 31) >> 2) ^ (row & 7)) << 4) + (size_t)(k & 3) * 4;
}

// Operand images of plane_analysis_tma_kernel.  E rows n < N1: tf32 hi of (n = 2l: cos, 2l+1: -sin)(2 pi l w / W), rows
// N1 + n: the lo parts; A2 rows grp * KHp + k (grp 0 cos hi, 1 sin hi, 2 cos lo, 3 sin lo of 2 pi f_k h / H) over the
// permuted row index K' = j * HGp + hh with h = g * hh + j.
void fill_analysis_tma_images(float* base, const Shape& s, const TwLayout& t) {
  const AtGeom a = at_geom(s);
  if (!a.ok) return;
  const double two_pi = 6.283185307179586476925286766559;
  const int ER = 2 * a.N1;
  for (int j = 0; j < a.g; ++j) {                 // one image per row class: column kc of the landed tile is w = kc - shift[j]
    float* e1 = base + t.at_e1 + (size_t)j * a.nchunks * ER * 32;
    for (int n = 0; n < a.N1; ++n)
      for (int kc = 0; kc < a.nchunks * 32; ++kc) {
        const int l = n / 2, w = kc - a.shift[j];
        double v = 0.0;
        if (l < s.MW && w >= 0 && w < s.W) {
          const double ang = two_pi * (double)(((long long)l * w) % s.W) / (double)s.W;
          v = (n & 1) ? -sin(ang) : cos(ang);
        }
        const float h = tf32_mask((float)v);
        e1[sw128_off(n, kc, ER) / 4] = h;
        e1[sw128_off(a.N1 + n, kc, ER) / 4] = (float)(v - (double)h);
      }
  }
  float* a2 = base + t.at_a2;
  const int AR = 4 * a.KHp;
  for (int k = 0; k < a.KHp; ++k)
    for (int kp = 0; kp < a.K2p; ++kp) {
      const int j = kp / a.HGp, hh = kp - j * a.HGp;
      double c = 0.0, sn = 0.0;
      if (k < s.KH && j < a.g && hh < a.HG) {
        const long long hrow = (long long)a.g * hh + j;
        const long long f = (k < s.mH) ? k : (long long)s.H - 2 * s.mH + k;
        const double ang = two_pi * (double)((f * hrow) % s.H) / (double)s.H;
        c = cos(ang);
        sn = sin(ang);
      }
      const float ch = tf32_mask((float)c), sh = tf32_mask((float)sn);
      a2[sw128_off(0 * a.KHp + k, kp, AR) / 4] = ch;
      a2[sw128_off(1 * a.KHp + k, kp, AR) / 4] = sh;
      a2[sw128_off(2 * a.KHp + k, kp, AR) / 4] = (float)(c - (double)ch);
      a2[sw128_off(3 * a.KHp + k, kp, AR) / 4] = (float)(sn - (double)sh);
    }
}

struct TwPtrs { const float2* last; const float2* mid; const float2* outer; const float2* outer2; const float* cl; const float* synth_ae;
                const float* an_e1; const float* an_a2; const float* at_e1; const float* at_a2; const float* rs_a; };

TwPtrs tw_ptrs(const Shape& s, const void* twiddles) {
  const float* base = static_cast<const float*>(twiddles);
  const TwLayout t = tw_layout(s);
  TwPtrs p;
  p.last = reinterpret_cast<const float2*>(base + t.last);
  p.mid = reinterpret_cast<const float2*>(base + t.mid);
  p.outer = reinterpret_cast<const float2*>(base + t.outer);
  p.outer2 = reinterpret_cast<const float2*>(base + t.outer2);
  p.cl = base + t.cl;
  p.synth_ae = base + t.synth_ae;
  p.an_e1 = base + t.an_e1;
  p.an_a2 = base + t.an_a2;
  p.at_e1 = base + t.at_e1;
  p.at_a2 = base + t.at_a2;
  p.rs_a = base + t.rs_a;
  return p;
}

// Process-wide switches, read from the environment ONCE (first use) and again only on deno_reload_env():
//   DENO_SPECTRAL_PATH = auto (default) | simt | tc : which kernels serve the stages both paths implement
//   DENO_MIX_TILE      = 0: never the tiled mix kernels, 2: tiled forward mix too      (experiments)
//   DENO_MIX_TM        = 1 | 2 | 4: modes per block of the tiled mix kernels            (experiments)
//   DENO_SYNTH_WS      = 1: warp-specialised synthesis kernel instead of the bulk one   (experiments)
enum class PathPref { Auto, Simt, Tc };
struct Options {
  PathPref path = PathPref::Auto;
  int mix_tile = 1;      // 0 / 1 (default) / 2
  int mix_tm = 0;        // 0 = heuristic
  bool synth_ws = false;
  bool wgrad_tma = true; // DENO_WGRAD_TMA=0: the register-staged bypass_wgrad_tc kernel instead of the TMA one
  bool mix_tc = true;    // DENO_MIX_TC=0: the SIMT mode-mix kernels also for width >= 64
  int at_dbg = 0;        // DENO_AT_DBG: bring-up switches of the TMA analysis kernel (see AnalysisTmaParams::dbg)
  int an_ring = 1;       // DENO_AN_RING: 0 = one staging buffer (round-1 layout), 1 = ring + staged act'(z) where it fits, 2 = ring, z by registers
  int mix_lane = 1;      // DENO_MIX_LANE: 0 = never, 1 (default) = mode_mix_lane (lane = mode, batch-tiled, staged spectrum) instead of the tensor-core mix for width >= 128, 2 = for every width
  int an_ablate = 0;     // DENO_AN_ABLATE: timing-only ablation bits of plane_analysis_tc (results wrong; see AnalysisTcParams::ablate)
  bool rowsynth_tc = true;  // DENO_ROWSYNTH_TC=0: the SIMT row_synthesis kernel instead of the tensor-core one
  bool an_tma = false;   // DENO_ANALYSIS_TMA=1: the TMA analysis kernel instead of the bulk-copy + SIMT-transform one (not yet green on hardware)
};
Options read_options() {
  Options o;
  if (const char* e = getenv("DENO_SPECTRAL_PATH")) {
    if (strcmp(e, "simt") == 0) o.path = PathPref::Simt;
    else if (strcmp(e, "tc") == 0) o.path = PathPref::Tc;
  }
  if (const char* e = getenv("DENO_MIX_TILE")) { if (e[0] == '0') o.mix_tile = 0; else if (e[0] == '2') o.mix_tile = 2; }
  if (const char* e = getenv("DENO_MIX_TM")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) o.mix_tm = v; }
  if (const char* e = getenv("DENO_SYNTH_WS")) o.synth_ws = e[0] == '1';
  if (const char* e = getenv("DENO_WGRAD_TMA")) o.wgrad_tma = e[0] != '0';
  if (const char* e = getenv("DENO_ANALYSIS_TMA")) o.an_tma = e[0] == '1';
  if (const char* e = getenv("DENO_AT_DBG")) o.at_dbg = atoi(e);
  if (const char* e = getenv("DENO_AN_RING")) o.an_ring = atoi(e);
  if (const char* e = getenv("DENO_ROWSYNTH_TC")) o.rowsynth_tc = e[0] != '0';
  if (const char* e = getenv("DENO_MIX_LANE")) o.mix_lane = atoi(e);
  if (const char* e = getenv("DENO_AN_ABLATE")) o.an_ablate = atoi(e);
  if (const char* e = getenv("DENO_MIX_TC")) o.mix_tc = e[0] != '0';
  return o;
}
std::mutex g_opt_mutex;
Options g_options;
bool g_options_read = false;
Options options() {
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  if (!g_options_read) { g_options = read_options(); g_options_read = true; }
  return g_options;
}
PathPref path_pref() { return options().path; }

#ifndef DENO_EMULATE
// TMA + SWIZZLE_128B analysis (spectral_tma.cuh); -1 when the shape / alignment is outside the kernel's coverage.
int try_launch_analysis_tma(const Shape& s, int channels, const float* u, const float* z, float* gz_out, float2* out,
                            const TwPtrs& tw, float scale, int herm, cudaStream_t st) {
  if (!options().an_tma) return -1;
  const AtGeom a = at_geom(s);
  if (!a.ok) return -1;
  const bool gz = (z != nullptr);
  if ((reinterpret_cast<uintptr_t>(u) & 15) != 0) return -1;
  if (gz && (((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gz_out)) & 15) != 0)) return -1;
  tc::AnalysisTmaParams p;
  memset(&p, 0, sizeof(p));
  p.H = s.H; p.W = s.W; p.g = a.g; p.HG = a.HG; p.HGp = a.HGp;
  p.KH = s.KH; p.KHp = a.KHp; p.MW = s.MW; p.N1 = a.N1;
  p.C = channels; p.X = s.X;
  const long long NP = (long long)s.B * channels * s.X;
  if (NP * a.HG >= (1LL << 31)) return -1;
  p.NP = (int)NP;
  p.nchunks = a.nchunks; p.K2p = a.K2p;
  // planes per tile: two when that fits (stage-1 MMAs cost max(~75, N / 2) cycles whatever N is), one for the gz variant
  // (three activation streams per plane: bandwidth-bound long before the tensor pipe) and when two do not fit
  int bestP = 0, best_nsh = 0, best_nsl = 0;
  for (int P = gz ? 1 : 2; P >= 1 && bestP == 0; --P) {
    p.P = P; p.N = a.g * P * a.HGp; p.N2 = 2 * P * a.N1;
    if (P * a.HGp > 256 || 2 * p.N + p.N2 > 512) continue;   // per-class stage-1 N = P * HGp
    for (int nsh = 4; nsh >= 2 && bestP == 0; --nsh) {
      p.nsh = nsh; p.nsl = 2;
      if (tc::analysis_tma_smem(p, gz).total + 1024 <= kMaxSmemBytes) { bestP = P; best_nsh = nsh; best_nsl = 2; }
    }
  }
  if (bestP == 0) return -1;
  p.P = bestP; p.N = a.g * bestP * a.HGp; p.N2 = 2 * bestP * a.N1; p.nsh = best_nsh; p.nsl = best_nsl;
  p.ntiles = (int)((NP + bestP - 1) / bestP);
  p.e1 = tw.at_e1; p.a2 = tw.at_a2; p.cl = tw.cl; p.scale = scale; p.herm = herm; p.out = out;
  p.u_base = u; p.z_base = z;
  p.l2_prefetch = ((long long)s.H * s.W) % 4 == 0 ? 1 : 0;
  p.dbg = options().at_dbg;
  if (p.dbg & 1) p.l2_prefetch = 0;
  tc::AnalysisTmaMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int j = 0; j < a.g; ++j) {
    const long long off_bytes = (long long)j * s.W * 4;
    const long long base_off = off_bytes & ~15LL;
    p.shift[j] = (int)((off_bytes - base_off) / 4);
    if (p.shift[j] != a.shift[j]) return -1;
    // extent rounded up to 16 bytes: a TMA STORE clips in 16-byte units (an extent of 94 floats let the class-0 store
    // zero the first two floats of the next row, gpurun r2i); the extra floats are the next row's head - loaded,
    // multiplied and stored back with the very values that row's own class writes
    const unsigned long long inner = (unsigned long long)round_up(s.W + p.shift[j], 4);
    const unsigned long long outer = (unsigned long long)NP * a.HG;
    const unsigned long long row_bytes = (unsigned long long)a.g * s.W * 4;
    if (!make_tmap_2d(&maps.u[j], reinterpret_cast<const char*>(u) + base_off, inner, outer, row_bytes, (unsigned)a.HG)) return -1;
    if (gz) {
      if (!make_tmap_2d(&maps.z[j], reinterpret_cast<const char*>(z) + base_off, inner, outer, row_bytes, (unsigned)a.HG)) return -1;
      if (!make_tmap_2d(&maps.gz[j], reinterpret_cast<const char*>(gz_out) + base_off, inner, outer, row_bytes, (unsigned)a.HG)) return -1;
    }
  }
  const size_t smem = (size_t)tc::analysis_tma_smem(p, gz).total + 1024;
  const int grid = p.ntiles < 148 ? p.ntiles : 148;
  int rc;
  if (gz) {
    rc = allow_smem(tc::plane_analysis_tma_kernel<true>, smem, "plane_analysis_tma");
    if (rc != DENO_OK) return rc;
    prof_before("plane_analysis_gz_tma", st);
    DENO_LAUNCH((tc::plane_analysis_tma_kernel<true>), dim3(grid), dim3(tc::kAtThreads), smem, st, maps, p);
  } else {
    rc = allow_smem(tc::plane_analysis_tma_kernel<false>, smem, "plane_analysis_tma");
    if (rc != DENO_OK) return rc;
    prof_before("plane_analysis_tma", st);
    DENO_LAUNCH((tc::plane_analysis_tma_kernel<false>), dim3(grid), dim3(tc::kAtThreads), smem, st, maps, p);
  }
  return check_launch("plane_analysis_tma", st);
}

// Tensor-core analysis; -1 when the shape is outside the kernel's coverage.
int try_launch_analysis_tc(const Shape& s, int channels, const float* u, const float* z, float* gz_out, float2* out,
                           const TwPtrs& tw, float scale, int herm, cudaStream_t st) {
  if (s.ndim < 2) return -1;
  const TwLayout tl = tw_layout(s);
  if (tl.an_nsub == 0) return -1;
  const bool gz = (z != nullptr);
  const tc::AnalysisTcSmem sm = tc::analysis_tc_smem(tl.an_HS, s.W, s.KH, tl.N1, tl.Wp, tl.Hp, tl.KRp, tl.an_nsub);
  if (sm.total > kMaxSmemBytes - 1024 || (long long)sm.total != an_tc_smem_total(tl.an_HS, s.W, s.KH, tl.N1, tl.Wp, tl.KRp, tl.an_nsub))
    return -1;
  tc::AnalysisTcParams p;
  p.g = make_geom(s, channels);
  p.u = u; p.z = z; p.gz_out = gz_out; p.out = out;
  p.e1 = tw.an_e1; p.a2 = tw.an_a2; p.cl = tw.cl;
  p.scale = scale; p.herm = herm; p.act = s.act;
  p.nplanes = s.B * s.X * channels;
  p.N1 = tl.N1; p.Wp = tl.Wp; p.Hp = tl.Hp; p.KS0 = tl.KS0; p.KRp = tl.KRp;
  p.nsub = tl.an_nsub; p.HS = tl.an_HS;
  // Staging ring: what fits next to the operand images.  Wanted: two slots (slab j + 1 lands while slab j is transformed;
  // with one slot its copy only starts when slab j has been consumed and ~0.9 k cycles of it are exposed per plane) and,
  // in the gz variant, act'(z) staged next to gy (the per-thread global loads of z were another ~1.5 k exposed cycles).
  // Tight fits (the 94 x 94 Darcy plane) put the epilogue's exchange buffer inside the B2 image and drop the
  // descriptor over-read pad, which the ring itself then covers.
  const int HWs = tl.an_HS * s.W, hlast = s.H - (tl.an_nsub - 1) * tl.an_HS;
  const bool bulk = ((long long)s.H * s.W) % 4 == 0 && (HWs % 4) == 0 && (hlast * s.W) % 4 == 0;
  const int stage_bytes = round_up((tl.an_nsub == 1 ? s.H : tl.an_HS) * s.W, 4) * 4;   // the rows that exist (HS is H rounded up to 8)
  const int limit = kMaxSmemBytes - 64;        // the barrier words (appended below) are the only other shared memory of the kernel
  int smem_total = sm.total;
  p.nslots = 1; p.zstage = 0; p.stage_off = sm.stage; p.xch_off = sm.xch; p.xch_alias = 0;
  p.ablate = options().an_ablate;
  if (bulk && options().an_ring) {
    const int zs = (gz && options().an_ring != 2) ? 1 : 0;
    const int want[3][2] = {{2, zs}, {zs ? 1 : 2, zs}, {1, 0}};   // (slots, z staged), best first
    for (int c = 0; c < 3; ++c) {
      const int bufs = want[c][0] * (1 + want[c][1]);
      if (bufs == 1) break;
      const int plain = sm.stage + bufs * stage_bytes + 4096;
      const int tight = sm.xch + bufs * stage_bytes;               // exchange buffer inside B2, no over-read pad (the ring covers it)
      if (plain <= limit) {
        p.nslots = want[c][0]; p.zstage = want[c][1]; smem_total = plain;
        break;
      }
      if (tight <= limit && bufs * stage_bytes >= 4096 && (sm.stage - sm.xch) <= sm.b2_bytes) {
        p.nslots = want[c][0]; p.zstage = want[c][1]; smem_total = tight;
        p.xch_alias = 1; p.xch_off = sm.b2; p.stage_off = sm.xch;
        break;
      }
    }
  }
  p.sync_off = round_up(smem_total, 16);
  smem_total = p.sync_off + 64;
  if (getenv("DENO_AN_DEBUG") != nullptr)
    fprintf(stderr, "[deno] analysis_tc H=%d W=%d gz=%d: base smem %d, stage %d, limit %d -> nslots %d zstage %d alias %d smem %d\n",
            s.H, s.W, (int)gz, sm.total, stage_bytes, limit, p.nslots, p.zstage, p.xch_alias, smem_total);
  const int grid = p.nplanes < 148 ? p.nplanes : 148;
  int rc;
  const int mode = !bulk ? 0 : (p.zstage ? (p.nslots == 2 ? 4 : 3) : (p.nslots == 2 ? 2 : 1));
#define DENO_AN_CASE6(GZ, SLABS, EVENW, MODEV, N1V, LABEL)                                                     \
  {                                                                                                            \
    rc = allow_smem(tc::plane_analysis_tc_kernel<GZ, SLABS, EVENW, MODEV, N1V>, (size_t)smem_total, "plane_analysis_tc");  \
    if (rc != DENO_OK) return rc;                                                                              \
    prof_before(LABEL, st);                                                                                    \
    DENO_LAUNCH((tc::plane_analysis_tc_kernel<GZ, SLABS, EVENW, MODEV, N1V>), dim3(grid), dim3(tc::kAnThreads), (size_t)smem_total, st, p); \
  }
#define DENO_AN_CASE5(GZ, SLABS, EVENW, MODEV, LABEL)                                                          \
  {                                                                                                            \
    DENO_AN_CASE6(GZ, SLABS, EVENW, MODEV, 0, LABEL)                                                           \
  }
#define DENO_AN_CASE_FWD(SLABS, EVENW)                                                                         \
  {                                                                                                            \
    if (mode == 0) DENO_AN_CASE5(false, SLABS, EVENW, 0, "plane_analysis_tc")                                  \
    else if (mode == 1) DENO_AN_CASE5(false, SLABS, EVENW, 1, "plane_analysis_tc")                             \
    else DENO_AN_CASE5(false, SLABS, EVENW, 2, "plane_analysis_tc")                                            \
  }
#define DENO_AN_CASE_GZ(SLABS, EVENW)                                                                          \
  {                                                                                                            \
    if (mode == 0) DENO_AN_CASE5(true, SLABS, EVENW, 0, "plane_analysis_gz_tc")                                \
    else if (mode == 1) DENO_AN_CASE5(true, SLABS, EVENW, 1, "plane_analysis_gz_tc")                           \
    else if (mode == 2) DENO_AN_CASE5(true, SLABS, EVENW, 2, "plane_analysis_gz_tc")                           \
    else if (mode == 3) DENO_AN_CASE5(true, SLABS, EVENW, 3, "plane_analysis_gz_tc")                           \
    else DENO_AN_CASE5(true, SLABS, EVENW, 4, "plane_analysis_gz_tc")                                          \
  }
  const bool slabs = p.nsub > 1;
  const bool evenw = (s.W % 2) == 0;
  if (gz) {
    if (slabs && evenw) DENO_AN_CASE_GZ(true, true)
    else if (slabs) DENO_AN_CASE_GZ(true, false)
    else if (evenw) DENO_AN_CASE_GZ(false, true)
    else DENO_AN_CASE_GZ(false, false)
  } else {
    if (slabs && evenw) DENO_AN_CASE_FWD(true, true)
    else if (slabs) DENO_AN_CASE_FWD(true, false)
    else if (evenw) DENO_AN_CASE_FWD(false, true)
    else DENO_AN_CASE_FWD(false, false)
  }
#undef DENO_AN_CASE_GZ
#undef DENO_AN_CASE_FWD
#undef DENO_AN_CASE5
#undef DENO_AN_CASE6
  return check_launch("plane_analysis_tc", st);
}
#endif

// x / (gy, z) -> packed plane spectra [B*X][C][KH][MW]
int launch_analysis(const Shape& s, int channels, const float* u, const float* z, float* gz_out, float2* out,
                    const TwPtrs& tw, float scale, int herm, cudaStream_t st) {
#ifndef DENO_EMULATE
  {
    const PathPref pref = s.simt_only ? PathPref::Simt : path_pref();
    if (pref != PathPref::Simt) {
      int trc = try_launch_analysis_tma(s, channels, u, z, gz_out, out, tw, scale, herm, st);
      if (trc >= 0) return trc;
      trc = try_launch_analysis_tc(s, channels, u, z, gz_out, out, tw, scale, herm, st);
      if (trc >= 0) return trc;
    }
  }
#endif
  if (s.ndim == 1 && s.MWp <= kLineThreads) {            // 1-D layers: lines, not planes (line_analysis_kernel)
    LineAnalysisParams lp;
    lp.g = make_geom(s, channels);
    lp.u = u; lp.z = z; lp.gz_out = gz_out; lp.out = out;
    lp.twW = tw.last; lp.cl = tw.cl; lp.scale = scale; lp.herm = herm;
    lp.nlines = s.B * s.X * channels;
    lp.MWp = s.MWp; lp.NS = kLineThreads / s.MWp;
    const size_t lsm = line_analysis_smem(s.W, lp.MWp, lp.NS);
    if (lsm <= (size_t)kMaxSmemBytes) {
      const int lgrid = (lp.nlines + kLineRows - 1) / kLineRows;
      int lrc;
      if (z != nullptr) {
        lrc = allow_smem(line_analysis_kernel<true>, lsm, "line_analysis");
        if (lrc != DENO_OK) return lrc;
        prof_before("line_analysis_gz", st);
        DENO_LAUNCH(line_analysis_kernel<true>, dim3(lgrid), dim3(kLineThreads), lsm, st, lp);
      } else {
        lrc = allow_smem(line_analysis_kernel<false>, lsm, "line_analysis");
        if (lrc != DENO_OK) return lrc;
        prof_before("line_analysis", st);
        DENO_LAUNCH(line_analysis_kernel<false>, dim3(lgrid), dim3(kLineThreads), lsm, st, lp);
      }
      return check_launch("line_analysis", st);
    }
  }
  AnalysisParams p;
  p.g = make_geom(s, channels);
  AnalysisCfg c;
  int rc = analysis_cfg(p.g, c);
  if (rc != DENO_OK) return rc;
  p.u = u; p.z = z; p.gz_out = gz_out; p.out = out;
  p.twW = tw.last; p.twH = tw.mid; p.cl = tw.cl;
  p.scale = scale; p.herm = herm; p.act = s.act;
  p.nplanes = s.B * s.X * channels;
  p.G = c.G; p.RT = c.RT; p.RTP = c.RTP; p.YC = c.YC; p.RS = c.RS; p.CS = c.CS;
  const int grid = (p.nplanes + c.G - 1) / c.G;
  if (z != nullptr) {
    rc = allow_smem(plane_analysis_kernel<true>, c.smem, "plane_analysis");
    if (rc != DENO_OK) return rc;
    prof_before("plane_analysis_gz", st);
    DENO_LAUNCH(plane_analysis_kernel<true>, dim3(grid), dim3(c.threads), c.smem, st, p);
  } else {
    rc = allow_smem(plane_analysis_kernel<false>, c.smem, "plane_analysis");
    if (rc != DENO_OK) return rc;
    prof_before("plane_analysis", st);
    DENO_LAUNCH(plane_analysis_kernel<false>, dim3(grid), dim3(c.threads), c.smem, st, p);
  }
  return check_launch("plane_analysis", st);
}

// One outer-axis pass: analysis  in [Bp][X][C][Q] -> out [Bp][C][KX][Q],  synthesis the other way round.
int launch_outer_axis(int Bp, int channels, int X, int KX, int Q, const float2* twX, bool analysis, const float2* in,
                      float2* out, cudaStream_t st) {
  OuterParams p;
  p.in = in; p.out = out; p.twX = twX;
  p.B = Bp; p.C = channels; p.X = X; p.KX = KX; p.Q = Q;
  const int KX4 = round_up(KX, 4);
  const long long units = analysis ? (long long)Bp * channels * (KX4 / 4) * Q
                                   : (long long)Bp * channels * ((X + 3) / 4) * Q;
  long long blocks = (units + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  prof_before(analysis ? "outer_analysis" : "outer_synthesis", st);
  if (analysis) {
    DENO_LAUNCH(outer_analysis_kernel, dim3((unsigned)blocks), dim3(256), 0, st, p);
  } else {
    DENO_LAUNCH(outer_synthesis_kernel, dim3((unsigned)blocks), dim3(256), 0, st, p);
  }
  return check_launch(analysis ? "outer_analysis" : "outer_synthesis", st);
}

int launch_outer(const Shape& s, int channels, bool analysis, const float2* in, float2* out, const TwPtrs& tw,
                 cudaStream_t st) {
  return launch_outer_axis(s.B, channels, s.X, s.KX, s.Q, tw.outer, analysis, in, out, st);
}

// 4-D: the two outer axes, one pass each.  Analysis: plane spectra [B*X1*X2][C][Q] -> (contract y) `mid` [B*X1][C][KX2][Q]
// -> (contract x; the (ky, q) pair is the inner index) packed [B][C][KX1][KX2*Q].  Synthesis runs the two passes backwards.
int launch_outer4(const Shape& s, int channels, bool analysis, const float2* in, float2* mid, float2* out, const TwPtrs& tw,
                  cudaStream_t st) {
  int rc;
  if (analysis) {
    if ((rc = launch_outer_axis(s.B * s.X1, channels, s.X2, s.KX2, s.Q, tw.outer2, true, in, mid, st))) return rc;
    return launch_outer_axis(s.B, channels, s.X1, s.KX1, s.KX2 * s.Q, tw.outer, true, mid, out, st);
  }
  if ((rc = launch_outer_axis(s.B, channels, s.X1, s.KX1, s.KX2 * s.Q, tw.outer, false, in, mid, st))) return rc;
  return launch_outer_axis(s.B * s.X1, channels, s.X2, s.KX2, s.Q, tw.outer2, false, mid, out, st);
}

// Tiled per-mode products (mode_mix_tile_kernel / mode_mix_bwd_tile_kernel): picks the modes per block so that the grid
// is about one block per SM and the staged operands fit shared memory; false when even one mode per block does not fit.
bool mix_tile_cfg(const Shape& s, bool backward, MixTileParams& p, size_t& smem_bytes, int& grid) {
  const Options opt = options();
  if (opt.mix_tile == 0) return false;
  p.md = make_decode(s);
  p.B = s.B; p.CI = s.CI; p.CO = s.CO;
  int TM = 4;
  while (TM > 1 && (s.M + TM - 1) / TM < 140) TM /= 2;
  if (opt.mix_tm != 0) TM = opt.mix_tm;   // experiments
  while (TM > 1 && (size_t)mix_tile_smem(s.B, s.CI, s.CO, TM, backward).total * 8 > (size_t)kMaxSmemBytes - 2048) TM /= 2;
  if ((size_t)mix_tile_smem(s.B, s.CI, s.CO, TM, backward).total * 8 > (size_t)kMaxSmemBytes - 2048) return false;
  p.TM = TM;
  const int rmax = s.CI > s.CO ? s.CI : s.CO;
  const int slots = kMixTileThreads / TM, nr2 = (rmax + 1) / 2;
  const int groups = slots / nr2 > 0 ? slots / nr2 : 1;
  int bper = (s.B + groups - 1) / groups;
  if (bper > kMixTileB) bper = kMixTileB;
  if (bper < 1) bper = 1;
  p.bper = bper;
  int cit = 4;
  while (cit > 1 && ((s.CI + cit - 1) / cit) * ((s.CO + 3) / 4) * TM < kMixTileThreads) cit /= 2;
  p.cit = cit;
  smem_bytes = (size_t)mix_tile_smem(s.B, s.CI, s.CO, TM, backward).total * 8 + 64;   // + the decoded weight offsets
  grid = (s.M + TM - 1) / TM;
  return true;
}

#ifndef DENO_EMULATE
// Tensor-core channel mix (spectral_mix_tc.cuh): width >= 64 on both sides (where the per-mode product is a real dense
// GEMM, north_star item 2), channels in tiles of 64, at most 64 batch entries (the GEMM N), enough modes to fill the
// machine.  Narrower layers keep the SIMT kernels.
bool mix_tc_eligible(const Shape& s) {
  if (path_pref() == PathPref::Simt || !options().mix_tc) return false;
  if (s.CI < 64 || s.CO < 64 || (s.CI % 64) != 0 || (s.CO % 64) != 0) return false;
  if (s.B > 64 || s.M < 64) return false;
  return true;
}
int mix_tc_modes_per_cta(const Shape& s, int ckp, int bp) {   // 2 when pairs of adjacent modes share 16-byte loads and fit
  const int mlast = s.ndim == 1 ? s.mW : s.MW;
  if ((mlast % 2) != 0 || (s.M % 2) != 0) return 1;
  return tc::mix_tc_smem(2, ckp, bp).total <= kMaxSmemBytes ? 2 : 1;
}

int launch_mix_tc(const Shape& s, bool backward, const float2* in, float2* out, const void* const* w_blocks, cudaStream_t st) {
  tc::MixTcParams p;
  memset(&p, 0, sizeof(p));
  p.in = in; p.out = out;
  p.md = make_decode(s);
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < 4; ++j) p.w[j] = static_cast<const float2*>(w_blocks[j < nblk ? j : 0]);
  p.B = s.B;
  if (!backward) {
    p.CK = s.CI; p.CR = s.CO;
    p.w_sck = (long long)s.CO * p.md.Mblk; p.w_scr = p.md.Mblk;
    p.conj_w = 0;
  } else {
    p.CK = s.CO; p.CR = s.CI;
    p.w_sck = p.md.Mblk; p.w_scr = (long long)s.CO * p.md.Mblk;
    p.conj_w = 1;
  }
  p.Bp = round_up(s.B, 8); p.CKp = round_up(p.CK, 8);
  p.TM = mix_tc_modes_per_cta(s, p.CKp, p.Bp);
  p.ntile = (p.CR + tc::kMixTcTile - 1) / tc::kMixTcTile;
  const size_t smem = (size_t)tc::mix_tc_smem(p.TM, p.CKp, p.Bp).total;
  if (smem > (size_t)kMaxSmemBytes) return -1;
  int rc = allow_smem(tc::mode_mix_tc_kernel, smem, "mode_mix_tc");
  if (rc != DENO_OK) return rc;
  const int grid = (s.M / p.TM) * p.ntile;
  prof_before(backward ? "mode_mix_bwd_tc" : "mode_mix_tc", st);
  DENO_LAUNCH((tc::mode_mix_tc_kernel), dim3(grid), dim3(tc::kMixTcThreads), smem, st, p);
  return check_launch("mode_mix_tc", st);
}

int launch_mix_wgrad_tc(const Shape& s, const float2* xhat, const float2* ghat, void* const* gw_blocks, float* glb_from_dc,
                        cudaStream_t st) {
  tc::MixWgradTcParams p;
  memset(&p, 0, sizeof(p));
  p.xhat = xhat; p.ghat = ghat;
  p.glb = glb_from_dc; p.s_total = (float)s.S;
  p.md = make_decode(s);
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < 4; ++j) p.gw[j] = static_cast<float2*>(gw_blocks[j < nblk ? j : 0]);
  p.B = s.B; p.CI = s.CI; p.CO = s.CO;
  p.Kp = round_up(s.B, 8);
  const int mlast = s.ndim == 1 ? s.mW : s.MW;
  p.TM = ((mlast % 2) == 0 && (s.M % 2) == 0 && tc::mix_wgrad_tc_smem(2, p.Kp).total <= kMaxSmemBytes) ? 2 : 1;
  p.nti = (s.CI + tc::kMixTcTile - 1) / tc::kMixTcTile;
  p.nto = (s.CO + tc::kMixTcTile - 1) / tc::kMixTcTile;
  const size_t smem = (size_t)tc::mix_wgrad_tc_smem(p.TM, p.Kp).total;
  if (smem > (size_t)kMaxSmemBytes) return -1;
  int rc = allow_smem(tc::mode_mix_wgrad_tc_kernel, smem, "mode_mix_wgrad_tc");
  if (rc != DENO_OK) return rc;
  const int grid = (s.M / p.TM) * p.nti * p.nto;
  prof_before("mode_mix_wgrad_tc", st);
  DENO_LAUNCH((tc::mode_mix_wgrad_tc_kernel), dim3(grid), dim3(tc::kMixTcThreads), smem, st, p);
  return check_launch("mode_mix_wgrad_tc", st);
}
#else
bool mix_tc_eligible(const Shape&) { return false; }
#endif

// lane = mode, batch-tiled product with the contracted spectrum staged in shared memory (mode_mix_lane_kernel)
int launch_mix_lane(const Shape& s, bool backward, const float2* in, float2* out, const void* const* w_blocks, cudaStream_t st) {
  MixParams p;
  p.in = in; p.out = out;
  p.md = make_decode(s);
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < 4; ++j) p.w[j] = static_cast<const float2*>(w_blocks[j < nblk ? j : 0]);
  p.B = s.B;
  if (!backward) {
    p.CI = s.CI; p.CO = s.CO;
    p.w_sci = (long long)s.CO * p.md.Mblk; p.w_sco = p.md.Mblk;
    p.conj_w = 0;
  } else {
    p.CI = s.CO; p.CO = s.CI;
    p.w_sci = p.md.Mblk; p.w_sco = (long long)s.CO * p.md.Mblk;
    p.conj_w = 1;
  }
  const size_t smem = mix_lane_smem_bytes(p.CI);
  int rc = allow_smem(mode_mix_lane_kernel, smem, "mode_mix_lane");
  if (rc != DENO_OK) return rc;
  dim3 block(32, 8);
  dim3 grid((s.M + 31) / 32, (p.CO + 7) / 8, (s.B + kMixLaneB - 1) / kMixLaneB);
  prof_before(backward ? "mode_mix_bwd_lane" : "mode_mix_lane", st);
  DENO_LAUNCH(mode_mix_lane_kernel, grid, block, smem, st, p);
  return check_launch("mode_mix_lane", st);
}

bool mix_lane_wanted(const Shape& s) {
  const int ml = options().mix_lane;
  if (ml == 2) return true;
  // measured (gpurun r3p / r3q): width 128 (C4b) layer 1865 -> 1822 us with the lane kernel, width 64 (C3) 283 -> 280
  // (a wash: the tensor-core mix stays there), width 32 (C2) 148 -> 154 (worse: the thread-per-output kernel has more warps)
  return ml == 1 && s.CI >= 128 && s.CO >= 128;
}

int launch_mix(const Shape& s, bool backward, const float2* in, float2* out, const void* const* w_blocks,
               cudaStream_t st) {
  if (mix_lane_wanted(s)) return launch_mix_lane(s, backward, in, out, w_blocks, st);
#ifndef DENO_EMULATE
  if (mix_tc_eligible(s)) {
    const int trc = launch_mix_tc(s, backward, in, out, w_blocks, st);
    if (trc >= 0) return trc;
  }
#endif
  MixParams p;
  p.in = in; p.out = out;
  p.md = make_decode(s);
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < 4; ++j) p.w[j] = static_cast<const float2*>(w_blocks[j < nblk ? j : 0]);
  p.B = s.B;
  if (!backward) {
    p.CI = s.CI; p.CO = s.CO;
    p.w_sci = (long long)s.CO * p.md.Mblk; p.w_sco = p.md.Mblk;
    p.conj_w = 0;
  } else {
    p.CI = s.CO; p.CO = s.CI;
    p.w_sci = p.md.Mblk; p.w_sco = (long long)s.CO * p.md.Mblk;
    p.conj_w = 1;
  }
  // The forward product alone is too little work per block for the tiled kernel (18.7 us vs 13.3 us for the
  // thread-per-output kernel at the Darcy shape, which fills every SM with warps); DENO_MIX_TILE=2 forces it.
  if (!backward && options().mix_tile == 2) {
    MixTileParams t;
    memset(&t, 0, sizeof(t));
    size_t tsm = 0;
    int tgrid = 0;
    if (mix_tile_cfg(s, false, t, tsm, tgrid)) {
      t.xhat = in; t.yhat = out;
      for (int j = 0; j < 4; ++j) t.w[j] = p.w[j];
      int rc = allow_smem(mode_mix_tile_kernel, tsm, "mode_mix_tile");
      if (rc != DENO_OK) return rc;
      prof_before("mode_mix", st);
      DENO_LAUNCH(mode_mix_tile_kernel, dim3(tgrid), dim3(kMixTileThreads), tsm, st, t);
      return check_launch("mode_mix_tile", st);
    }
  }
  dim3 block(32, 8);
  dim3 grid((s.M + 31) / 32, (p.CO + 7) / 8, (s.B + kMixBatchTile - 1) / kMixBatchTile);
  prof_before(backward ? "mode_mix_bwd" : "mode_mix", st);
  DENO_LAUNCH(mode_mix_kernel, grid, block, 0, st, p);
  return check_launch("mode_mix", st);
}

int launch_mix_wgrad(const Shape& s, const float2* xhat, const float2* ghat, void* const* gw_blocks, cudaStream_t st) {
  MixWgradParams p;
  p.xhat = xhat; p.ghat = ghat;
  p.md = make_decode(s);
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < 4; ++j) p.gw[j] = static_cast<float2*>(gw_blocks[j < nblk ? j : 0]);
  p.B = s.B; p.CI = s.CI; p.CO = s.CO;
  dim3 block(32, 8);
  dim3 grid((s.M + 31) / 32, ((s.CO + kMixWgradOut - 1) / kMixWgradOut + 7) / 8, s.CI);
  prof_before("mode_mix_wgrad", st);
  DENO_LAUNCH(mode_mix_wgrad_kernel, grid, block, 0, st, p);
  return check_launch("mode_mix_wgrad", st);
}

// Fused launch of the backward consumers of G^ (see mode_mix_bwd_fused_kernel).  Any role may be disabled.
int launch_mix_bwd_fused(const Shape& s, const float2* xhat, const float2* ghat, float2* gxhat, const void* const* w_blocks,
                         void* const* gw_blocks, float* glb_from_dc, cudaStream_t st,
                         const char* label = "mode_mix_bwd_fused") {
  const int nblk = num_weight_blocks(s);
  if (mix_lane_wanted(s) && gxhat != nullptr && w_blocks != nullptr) {   // G^x by the lane kernel, the other roles below
    const int lrc = launch_mix_lane(s, true, ghat, gxhat, w_blocks, st);
    if (lrc != DENO_OK) return lrc;
    gxhat = nullptr;
    if (gw_blocks == nullptr && glb_from_dc == nullptr) return DENO_OK;
  }
#ifndef DENO_EMULATE
  if (mix_tc_eligible(s)) {   // width >= 64: two tensor-core launches (+ the DC-mode bias gradient) instead of the fused SIMT one
    int rc = DENO_OK;
    if (gxhat != nullptr && w_blocks != nullptr) {
      rc = launch_mix_tc(s, true, ghat, gxhat, w_blocks, st);
      if (rc > 0) return rc;
    }
    bool bias_done = false;
    if (rc == DENO_OK && gw_blocks != nullptr) {
      rc = launch_mix_wgrad_tc(s, xhat, ghat, gw_blocks, glb_from_dc, st);
      if (rc > 0) return rc;
      bias_done = rc == DENO_OK;
    }
    if (rc == DENO_OK) {
      if (glb_from_dc != nullptr && !bias_done) {
        tc::BiasDcParams bp;
        bp.ghat = ghat; bp.glb = glb_from_dc; bp.B = s.B; bp.CO = s.CO; bp.M = s.M; bp.s_total = (float)s.S;
        prof_before("bias_from_dc", st);
        DENO_LAUNCH((tc::bias_from_dc_kernel), dim3((s.CO + 127) / 128), dim3(128), 0, st, bp);
        return check_launch("bias_from_dc", st);
      }
      return DENO_OK;
    }
    // rc < 0: a kernel declined before launching anything that cannot be redone below
  }
#endif
  {
    MixTileParams t;
    memset(&t, 0, sizeof(t));
    size_t tsm = 0;
    int tgrid = 0;
    if ((gxhat != nullptr || gw_blocks != nullptr || glb_from_dc != nullptr) && mix_tile_cfg(s, true, t, tsm, tgrid)) {
      t.xhat = xhat; t.ghat = ghat; t.gxhat = gxhat;
      for (int j = 0; j < 4; ++j) {
        t.w[j] = (gxhat != nullptr && w_blocks != nullptr) ? static_cast<const float2*>(w_blocks[j < nblk ? j : 0]) : nullptr;
        t.gw[j] = gw_blocks != nullptr ? static_cast<float2*>(gw_blocks[j < nblk ? j : 0]) : nullptr;
      }
      t.glb = glb_from_dc; t.s_total = (float)s.S;
      int rc = allow_smem(mode_mix_bwd_tile_kernel, tsm, "mode_mix_bwd_tile");
      if (rc != DENO_OK) return rc;
      prof_before(label, st);
      DENO_LAUNCH(mode_mix_bwd_tile_kernel, dim3(tgrid), dim3(kMixTileThreads), tsm, st, t);
      return check_launch("mode_mix_bwd_tile", st);
    }
  }
  MixBwdFusedParams f;
  memset(&f, 0, sizeof(f));
  f.mix.md = make_decode(s);
  f.wg.md = f.mix.md;
  if (gxhat != nullptr) {
    f.mix.in = ghat; f.mix.out = gxhat;
    for (int j = 0; j < 4; ++j) f.mix.w[j] = static_cast<const float2*>(w_blocks[j < nblk ? j : 0]);
    f.mix.B = s.B; f.mix.CI = s.CO; f.mix.CO = s.CI;
    f.mix.w_sci = f.mix.md.Mblk; f.mix.w_sco = (long long)s.CO * f.mix.md.Mblk;
    f.mix.conj_w = 1;
    f.ax = (s.M + 31) / 32; f.ay = (s.CI + 7) / 8; f.az = (s.B + kMixBatchTile - 1) / kMixBatchTile;
  }
  if (gw_blocks != nullptr) {
    f.wg.xhat = xhat; f.wg.ghat = ghat;
    for (int j = 0; j < 4; ++j) f.wg.gw[j] = static_cast<float2*>(gw_blocks[j < nblk ? j : 0]);
    f.wg.B = s.B; f.wg.CI = s.CI; f.wg.CO = s.CO;
    f.bx = (s.M + 31) / 32; f.by = ((s.CO + kMixWgradOut - 1) / kMixWgradOut + 7) / 8; f.bz = s.CI;
  }
  f.ghat = ghat; f.glb = glb_from_dc; f.B = s.B; f.CO = s.CO; f.M = s.M; f.s_total = (float)s.S;
  const int total = f.ax * f.ay * f.az + f.bx * f.by * f.bz + (glb_from_dc != nullptr ? 1 : 0);
  if (total == 0) return DENO_OK;
  prof_before(label, st);
  DENO_LAUNCH(mode_mix_bwd_fused_kernel, dim3(total), dim3(32, 8), 0, st, f);
  return check_launch("mode_mix_bwd_fused", st);
}

#ifndef DENO_EMULATE
// The warp-specialised synthesis variant (24 template instances, slower than the bulk kernel at every measured shape) is
// compiled only with -DDENO_WITH_SYNTH_WS (DENO_EXTRA_NVCC_FLAGS); without it DENO_SYNTH_WS=1 falls through to the bulk kernel.
#ifdef DENO_WITH_SYNTH_WS
// Tensor-core synthesis: returns DENO_OK after launching, or -1 when the shape is outside what the
// tcgen05 kernel covers (caller falls back to the SIMT kernel).
int try_launch_synthesis_ws(const Shape& s, SynthParams p, const TwPtrs& tw, int act, const char* label, cudaStream_t st) {
  const int CI = p.g.C, CO = p.CO;
  if (CI > 64 || CO > 128 || (CO % 32) != 0) return -1;   // CO / 4 columns per epilogue warp, multiples of 8
  const int CIP = CI <= 8 ? 8 : (CI <= 16 ? 16 : (CI <= 32 ? 32 : 64));
  const TwLayout tl = tw_layout(s);
  // stage A stages the packed spectrum through the A_X area: at least one output channel must fit
  if (2 * CIP * 128 * 4 - round_up(kSynthMaxRows * p.g.KH, 2) * 8 < p.g.KH * p.g.MW * 8) return -1;
  int cols = 32;
  while (cols < 6 * CO) cols *= 2;
  if (cols > 512) return -1;
  // One persistent, warp-specialised CTA per SM; pick the rows per unit that balance the units over 148 CTAs.
  int bestR = 0;
  double best_eff = -1.0;
  const int maxR = p.g.H > 1 ? (p.g.H < kSynthMaxRows ? p.g.H : kSynthMaxRows) : 1;
  for (int R = maxR; R >= 1; --R) {
    const size_t bytes = (size_t)tc::synth_ws_smem(CIP, CO, tl.KE, R).total;
    if (bytes > (size_t)kMaxSmemBytes - 2048) continue;
    const long long units = (long long)p.g.NBO * p.g.NBI * ((p.g.H + R - 1) / R) * tl.nchunks;
    const long long waves = (units + 147) / 148;
    double eff = (double)units / (double)(waves * 148);
    eff *= (1.0 - 0.3 / R);               // per-unit drain + stage A amortise over R rows
    if (eff > best_eff + 1e-9) { best_eff = eff; bestR = R; }
  }
  if (bestR == 0) return -1;
  tc::SynthTcParams q;
  p.R = bestR;
  p.PTS = bestR * p.g.W;
  p.tiles = (p.g.H + bestR - 1) / bestR;
  q.s = p;
  q.ae = tw.synth_ae;
  q.urow = nullptr;
  q.KE = tl.KE;
  q.nchunks = tl.nchunks;
  q.tmem_cols = cols;
  q.COT = CO; q.nco = 1; q.npass = 1;
  const size_t smem = (size_t)tc::synth_ws_smem(CIP, CO, tl.KE, bestR).total;
  const long long nunits = (long long)s.B * s.X * p.tiles * tl.nchunks;
  const int grid = (int)(nunits < 148 ? nunits : 148);
  int rc = DENO_OK;
#define DENO_TC_CASE2(ACT, CIPV)                                                                    \
  {                                                                                                 \
    rc = allow_smem(tc::plane_synthesis_ws_kernel<ACT, CIPV>, smem, "plane_synthesis_ws");          \
    if (rc != DENO_OK) return rc;                                                                   \
    prof_before(label, st);                                                                         \
    DENO_LAUNCH((tc::plane_synthesis_ws_kernel<ACT, CIPV>), dim3(grid), dim3(tc::kWsThreads), smem, st, q);    \
  }
#define DENO_TC_CASE(ACT)                                                                           \
  case ACT:                                                                                         \
    if (CIP == 8) DENO_TC_CASE2(ACT, 8)                                                             \
    else if (CIP == 16) DENO_TC_CASE2(ACT, 16)                                                      \
    else if (CIP == 32) DENO_TC_CASE2(ACT, 32)                                                      \
    else DENO_TC_CASE2(ACT, 64)                                                                     \
    break;
  switch (act) {
    DENO_TC_CASE(ACT_IDENTITY)
    DENO_TC_CASE(ACT_GELU)
    DENO_TC_CASE(ACT_SILU)
    DENO_TC_CASE(ACT_RELU)
    DENO_TC_CASE(ACT_TANH)
    DENO_TC_CASE(ACT_LEAKYRELU)
    default: return fail(DENO_E_INVALID, "unknown activation id");
  }
#undef DENO_TC_CASE
#undef DENO_TC_CASE2
  return check_launch("plane_synthesis_tc", st);
}

#endif  // DENO_WITH_SYNTH_WS

// Bulk-synchronous variant (two CTAs per SM, every warp does every phase).  Measured faster than the
// warp-specialised kernel at the Darcy shape (profiles/r1t_phase_probe.txt), so it is the default;
// DENO_SYNTH_WS=1 selects the warp-specialised kernel.
int try_launch_synthesis_bulk(const Shape& s, SynthParams p, const TwPtrs& tw, int act, const char* label, float* urow, cudaStream_t st) {
  const int CI = p.g.C, CO = p.CO;
  if (CI > 256 || CO > 256 || (CO % 32) != 0) return -1;
  // Wide layers (Airfoil / Pipe, width 128): the output channels are cut into tiles of 64 (one CTA each: 6 * COT TMEM
  // columns for the double-buffered split-pass accumulators) and the input channels go through the A_X buffer in K
  // passes of 64.  Up to 64 channels nothing changes (one tile, one pass).
  const int COT = (6 * CO <= 512) ? CO : 64;
  if (CO % COT != 0) return -1;
  const int nco = CO / COT;
  const bool tiled = nco > 1 || CI > 64;                    // tiled instances are built for K passes of 64 channels only
  const int CIP = tiled ? 64 : (CI <= 8 ? 8 : (CI <= 16 ? 16 : (CI <= 32 ? 32 : 64)));
  const int npass = (CI + CIP - 1) / CIP;
  const TwLayout tl = tw_layout(s);
  if (tiled && (urow == nullptr || p.g.H <= 1)) return -1;   // the in-kernel row-axis stage knows one tile / one pass only
  // stage A stages the packed spectrum through the A_X area: at least one output channel must fit
  if (!tiled && 2 * CIP * 128 * 4 - round_up(kSynthMaxRows * p.g.KH, 2) * 8 < p.g.KH * p.g.MW * 8) return -1;
  int bestR = 0;
  double best_eff = -1.0;
  const int maxR = p.g.H > 1 ? (p.g.H < kSynthMaxRows ? p.g.H : kSynthMaxRows) : 1;
  for (int R = maxR; R >= 1; --R) {
    const size_t bytes = (size_t)tc::synth_tc_smem(CIP, COT, tl.KE, R, npass).total;
    if (bytes > (size_t)kMaxSmemBytes) continue;
    int per_sm = (int)((228 * 1024) / (bytes + 2048));
    if (per_sm > 4) per_sm = 4;
    if (per_sm * 6 * COT > 512) per_sm = 512 / (6 * COT) > 0 ? 512 / (6 * COT) : 1;   // TMEM columns per SM
    if (per_sm < 1) per_sm = 1;
    const long long ctas = (long long)p.g.NBO * p.g.NBI * ((p.g.H + R - 1) / R) * tl.nchunks * nco;
    const long long slots = 148LL * per_sm;
    const long long waves = (ctas + slots - 1) / slots;
    double eff = (double)ctas / (double)(waves * slots);
    eff *= (1.0 - 0.25 / R);               // per-CTA setup (weights, twiddles, stage A) amortises over R rows
    if (eff > best_eff + 1e-9) { best_eff = eff; bestR = R; }
  }
  if (bestR == 0) return -1;
  tc::SynthTcParams q;
  p.R = bestR;
  p.PTS = bestR * p.g.W;
  p.tiles = (p.g.H + bestR - 1) / bestR;
  q.s = p;
  q.ae = tw.synth_ae;
  q.urow = nullptr;
  if (urow != nullptr && p.g.H > 1) {                   // row-axis synthesis once per layer instead of once per tile
    tc::RowSynthParams r;
    r.spec = p.spec; r.twH = p.twH; r.cl = p.cl; r.u = urow; r.scale = p.scale; r.herm = p.herm;
    r.H = p.g.H; r.KH = p.g.KH; r.KH4 = round_up(p.g.KH, 4); r.MW = p.g.MW; r.CO = CO; r.KE = tl.KE;
    r.Qp = tc::row_synth_pitch(r.KH, r.MW, r.KE);
    r.tiles = (r.H + tc::kRowSynthRows - 1) / tc::kRowSynthRows;
    r.ocb = CO;   // whole spectrum per block (8-channel blocks with staged 128-byte stores measured slower: 22 vs 16 us)
    while (r.ocb > 8 && tc::row_synth_smem_bytes(r.ocb, r.KH, r.MW, r.KE) > (size_t)kMaxSmemBytes) r.ocb = (r.ocb + 1) / 2;   // width 128: two halves
    r.ogroups = (CO + r.ocb - 1) / r.ocb;
    r.COT = COT; r.nco = nco;
    const size_t rs_smem = tc::row_synth_smem_bytes(r.ocb, r.KH, r.MW, r.KE);
    // tensor-core version: one GEMM per (batch entry, channel tile, G image column quads)
    bool rs_done = false;
    if (options().rowsynth_tc && tl.rs_ntm > 0 && (COT % 8) == 0 && (tl.KE % 4) == 0) {
      tc::RowSynthTcParams rt;
      rt.spec = p.spec; rt.ta = tw.rs_a; rt.cl = p.cl; rt.u = urow; rt.scale = p.scale; rt.herm = p.herm;
      rt.H = p.g.H; rt.KH = p.g.KH; rt.MW = p.g.MW; rt.CO = CO; rt.KE = tl.KE; rt.KP = tl.rs_KP;
      rt.COT = COT; rt.nco = nco; rt.ntm = tl.rs_ntm;
      const int nkq = tl.KE / 4;
      const long long units = (long long)p.g.NBO * p.g.NBI * nco;
      int G = 0, Gmin = 0;
      for (int g = nkq; g >= 1; --g) {                 // largest column group that still gives every SM a CTA
        if (nkq % g != 0 || g * 4 * COT > 256) continue;
        if (tl.rs_ntm * g * 4 * COT > 512) continue;   // TMEM columns
        if ((size_t)tc::row_synth_tc_smem(rt.KP, g * 4 * COT, rt.ntm).total > (size_t)kMaxSmemBytes - 256) continue;
        Gmin = g;
        if (G == 0 && units * (nkq / g) >= 148) G = g;
      }
      if (G == 0) G = Gmin;                            // small problems: as many CTAs as the shape allows
      if (G > 0) {
        rt.G = G; rt.ngrp = nkq / G;
        const size_t tsm = (size_t)tc::row_synth_tc_smem(rt.KP, G * 4 * COT, rt.ntm).total;
        int rrc = allow_smem(tc::row_synthesis_tc_kernel, tsm, "row_synthesis_tc");
        if (rrc != DENO_OK) return rrc;
        prof_before("row_synthesis_tc", st);
        DENO_LAUNCH((tc::row_synthesis_tc_kernel), dim3((unsigned)(units * rt.ngrp)), dim3(tc::kRsTcThreads), tsm, st, rt);
        rrc = check_launch("row_synthesis_tc", st);
        if (rrc != DENO_OK) return rrc;
        q.urow = urow;
        rs_done = true;
      }
    }
    if (!rs_done && rs_smem <= (size_t)kMaxSmemBytes) {
      int rrc = allow_smem(tc::row_synthesis_kernel, rs_smem, "row_synthesis");
      if (rrc != DENO_OK) return rrc;
      int threads = round_up(r.ocb * (tl.KE / 2), 32);
      if (threads > 384) threads = 384;
      prof_before("row_synthesis", st);
      DENO_LAUNCH((tc::row_synthesis_kernel), dim3((unsigned)(p.g.NBO * p.g.NBI * r.tiles * r.ogroups)), dim3(threads), rs_smem, st, r);
      rrc = check_launch("row_synthesis", st);
      if (rrc != DENO_OK) return rrc;
      q.urow = urow;
    }
  }
  if (tiled && q.urow == nullptr) return -1;
  q.KE = tl.KE;
  q.nchunks = tl.nchunks;
  q.COT = COT; q.nco = nco; q.npass = npass;
  int cols = 32;
  while (cols < 6 * COT) cols *= 2;
  if (cols > 512) return -1;
  q.tmem_cols = cols;
  const size_t smem = (size_t)tc::synth_tc_smem(CIP, COT, tl.KE, bestR, npass).total;
  const int grid = s.B * s.X * p.tiles * tl.nchunks * nco;
  int rc = DENO_OK;
#define DENO_TC_CASE5(ACT, CIPV, TILEDV, STAGEA, SQV)                                               \
  {                                                                                                 \
    rc = allow_smem(tc::plane_synthesis_tc_kernel<ACT, CIPV, TILEDV, STAGEA, SQV>, smem, "plane_synthesis_tc");  \
    if (rc != DENO_OK) return rc;                                                                   \
    prof_before(label, st);                                                                         \
    DENO_LAUNCH((tc::plane_synthesis_tc_kernel<ACT, CIPV, TILEDV, STAGEA, SQV>), dim3(grid), dim3(tc::kSynThreads), smem, st, q);       \
  }
#define DENO_TC_CASE4(ACT, CIPV, TILEDV, STAGEA)                                                    \
  {                                                                                                 \
    if (COT == CIPV) DENO_TC_CASE5(ACT, CIPV, TILEDV, STAGEA, true)                                 \
    else DENO_TC_CASE5(ACT, CIPV, TILEDV, STAGEA, false)                                            \
  }
#define DENO_TC_CASE3(ACT, CIPV, TILEDV)                                                            \
  {                                                                                                 \
    if (q.urow == nullptr) DENO_TC_CASE4(ACT, CIPV, false, true)                                    \
    else DENO_TC_CASE4(ACT, CIPV, TILEDV, false)                                                    \
  }
#define DENO_TC_CASE(ACT)                                                                           \
  case ACT:                                                                                         \
    if (tiled) DENO_TC_CASE4(ACT, 64, true, false)                                                  \
    else if (CIP == 8) DENO_TC_CASE3(ACT, 8, false)                                                 \
    else if (CIP == 16) DENO_TC_CASE3(ACT, 16, false)                                               \
    else if (CIP == 32) DENO_TC_CASE3(ACT, 32, false)                                               \
    else DENO_TC_CASE3(ACT, 64, false)                                                              \
    break;
  switch (act) {
    DENO_TC_CASE(ACT_IDENTITY)
    DENO_TC_CASE(ACT_GELU)
    DENO_TC_CASE(ACT_SILU)
    DENO_TC_CASE(ACT_RELU)
    DENO_TC_CASE(ACT_TANH)
    DENO_TC_CASE(ACT_LEAKYRELU)
    default: return fail(DENO_E_INVALID, "unknown activation id");
  }
#undef DENO_TC_CASE
#undef DENO_TC_CASE4
#undef DENO_TC_CASE5
#undef DENO_TC_CASE3
  return check_launch("plane_synthesis_tc", st);
}

int try_launch_synthesis_tc(const Shape& s, SynthParams p, const TwPtrs& tw, int act, const char* label, float* urow,
                            cudaStream_t st) {
#ifdef DENO_WITH_SYNTH_WS
  if (options().synth_ws) return try_launch_synthesis_ws(s, p, tw, act, label, st);
#endif
  return try_launch_synthesis_bulk(s, p, tw, act, label, urow, st);
}
#endif

// spec [B*X][CO][KH][MW] + bypass(u) -> y (, z)
int launch_synthesis(const Shape& s, int cin, int cout, const float2* spec, const float* u, const float* lw,
                     long long lw_so, long long lw_si, const float* bias, float* y, float* z, const TwPtrs& tw,
                     float scale, int herm, int act, float* urow, cudaStream_t st) {
  SynthParams p;
  p.g = make_geom(s, cin);
  p.CO = cout;
  SynthCfg c;
  int rc = synth_cfg(p.g, cout, c);
  if (rc != DENO_OK) return rc;
  p.sbo_out = (long long)cout * s.S;
  p.spec = spec; p.u = u; p.lw = lw; p.lw_so = lw_so; p.lw_si = lw_si; p.bias = bias;
  p.y = y; p.z = z;
  p.twW = tw.last; p.twH = tw.mid; p.cl = tw.cl;
  p.scale = scale; p.herm = herm; p.act = act;
  p.R = c.R; p.PTS = c.PTS; p.tiles = c.tiles;
  const int grid = s.B * s.X * c.tiles;
  const char* label = (bias != nullptr || act != DENO_ACT_IDENTITY || z != nullptr) ? "plane_synthesis" : "plane_synthesis_bwd";
#ifndef DENO_EMULATE
  const PathPref pref = s.simt_only ? PathPref::Simt : path_pref();
  if (pref != PathPref::Simt) {
    const bool is_fwd = (bias != nullptr || act != DENO_ACT_IDENTITY || z != nullptr);
    rc = try_launch_synthesis_tc(s, p, tw, act, is_fwd ? "plane_synthesis_tc" : "plane_synthesis_bwd_tc", urow, st);
    if (rc >= 0) return rc;
    if (pref == PathPref::Tc) return fail(DENO_E_UNSUPPORTED, "DENO_SPECTRAL_PATH=tc but the shape is outside the tcgen05 synthesis kernel");
  }
#endif
#define DENO_SYNTH_CASE(ACT)                                                                       \
  case ACT:                                                                                        \
    rc = allow_smem(plane_synthesis_kernel<ACT>, c.smem, "plane_synthesis");                       \
    if (rc != DENO_OK) return rc;                                                                  \
    prof_before(label, st);                                                                        \
    DENO_LAUNCH(plane_synthesis_kernel<ACT>, dim3(grid), dim3(c.threads), c.smem, st, p);          \
    break;
  switch (act) {
    DENO_SYNTH_CASE(ACT_IDENTITY)
    DENO_SYNTH_CASE(ACT_GELU)
    DENO_SYNTH_CASE(ACT_SILU)
    DENO_SYNTH_CASE(ACT_RELU)
    DENO_SYNTH_CASE(ACT_TANH)
    DENO_SYNTH_CASE(ACT_LEAKYRELU)
    default: return fail(DENO_E_INVALID, "unknown activation id");
  }
#undef DENO_SYNTH_CASE
  return check_launch("plane_synthesis", st);
}

#ifndef DENO_EMULATE
// Tensor-core bypass weight gradient; returns the number of partials written, -1 if not applicable, -2 on error.
int try_launch_bypass_wgrad_tc(const Shape& s, const float* gz, const float* x, float* partial, cudaStream_t st, int gz_tps = 0) {
  const int COs = round_up(s.CO, 8), CIs = round_up(s.CI, 8);
  if (COs > 128 || CIs > 128) return -1;
  int NS = 4;
  while (NS > 1 && (NS * COs > 128 || NS * CIs > 128)) NS /= 2;
  if ((NS * CIs) % 16 != 0) return -1;
  tc::WgradTcParams p;
  p.gx = make_geom(s, s.CI);
  p.CO = s.CO;
  p.sbo_g = (long long)s.CO * s.S;
  p.gz = gz; p.x = x; p.partial = partial;
  p.NS = NS; p.COs = COs; p.CIs = CIs;
  const int HW = s.H * s.W;
  p.tiles_per_plane = (HW + tc::kWgradKT - 1) / tc::kWgradKT;
  p.nbatch = s.B * s.X;
  p.ngroups = (p.nbatch + NS - 1) / NS;
  p.ntiles = p.ngroups * p.tiles_per_plane;
  p.vec4 = (HW % 4 == 0) ? 1 : 0;
  p.gz_tps = gz_tps;
  const tc::WgradTcSmem sm = tc::wgrad_tc_smem(NS, COs, CIs);
  if (sm.total > kMaxSmemBytes - 1024) return -1;
  const int need = (NS * (s.CO + s.CI) * (tc::kWgradKT / 4) + tc::kWgradLoaders - 1) / tc::kWgradLoaders;
  if (need > 8) return -1;
  int grid = 148;                         // one persistent CTA per SM (the accumulators take up to 512 TMEM columns)
  if (grid > p.ntiles) grid = p.ntiles;
  int rc;
#define DENO_WG_CASE(NITV)                                                                         \
  {                                                                                                \
    rc = allow_smem(tc::bypass_wgrad_tc_kernel<NITV>, (size_t)sm.total, "bypass_wgrad_tc");        \
    if (rc != DENO_OK) return -2;                                                                  \
    prof_before("bypass_wgrad_tc", st);                                                            \
    DENO_LAUNCH((tc::bypass_wgrad_tc_kernel<NITV>), dim3(grid), dim3(tc::kWgradThreads), (size_t)sm.total, st, p);          \
  }
  if (need <= 2) DENO_WG_CASE(2)
  else if (need <= 4) DENO_WG_CASE(4)
  else DENO_WG_CASE(8)
#undef DENO_WG_CASE
  if (check_launch("bypass_wgrad_tc", st) != DENO_OK) return -2;
  return grid * NS;
}
#endif

#ifndef DENO_EMULATE
// TMA + SWIZZLE_128B bypass weight gradient (spectral_tma.cuh); returns the number of partials written (one per CTA),
// -1 if the shape / alignment is outside the kernel, -2 on a CUDA error.
int try_launch_bypass_wgrad_tma(const Shape& s, const float* gz, const float* x, float* partial, cudaStream_t st, int gz_tps = 0) {
  if (!options().wgrad_tma) return -1;
  const int COs = round_up(s.CO, 8), CIs = round_up(s.CI, 8);
  if (COs > 128 || CIs > 128) return -1;
  int NS = 4;
  while (NS > 1 && (NS * COs > 128 || NS * CIs > 128)) NS /= 2;
  if ((NS * CIs) % 16 != 0) return -1;
  if ((s.S % 4) != 0) return -1;                          // rows of the [B*C][S] matrices must start 16-byte aligned
  if (COs != s.CO || CIs != s.CI) return -1;              // boxes never straddle the end of the tensor along its rows
  const int nbatch = s.B;                                  // S spans the whole (X, H, W) volume of a batch entry
  CUtensorMap tm_a, tm_b;
  if (gz_tps > 0) {
    if (!make_tmap_2d(&tm_a, gz, 128, (unsigned long long)nbatch * gz_tps * s.CO, 512, (unsigned)COs)) return -1;
  } else {
    if (!make_tmap_2d(&tm_a, gz, (unsigned long long)s.S, (unsigned long long)nbatch * s.CO, (unsigned long long)s.S * 4, (unsigned)COs)) return -1;
  }
  if (!make_tmap_2d(&tm_b, x, (unsigned long long)s.S, (unsigned long long)nbatch * s.CI, (unsigned long long)s.S * 4, (unsigned)CIs)) return -1;
  tc::WgradTmaParams p;
  p.partial = partial;
  p.CO = s.CO; p.CI = s.CI; p.COs = COs; p.CIs = CIs; p.NS = NS;
  p.nchunks = (int)((s.S + 31) / 32);
  p.ngroups = (nbatch + NS - 1) / NS;
  p.ntiles = p.ngroups * p.nchunks;
  p.a_rows_per_batch = s.CO; p.b_rows_per_batch = s.CI;
  p.a_tps = gz_tps;
  p.nbatch = nbatch;
  const int stage_bytes = tc::wgrad_tma_stage_bytes(NS * CIs);
  int nst = (kMaxSmemBytes - 2048) / stage_bytes;
  if (nst > 4) nst = 4;
  if (nst < 2) return -1;
  if ((size_t)NS * s.CO * s.CI * 4 > (size_t)nst * stage_bytes) return -1;   // epilogue staging lives in the stages
  p.nstages = nst;
  const size_t smem = (size_t)nst * stage_bytes + 1024;
  int grid = 148;
  if (grid > p.ntiles) grid = p.ntiles;
  if (allow_smem(tc::bypass_wgrad_tma_kernel, smem, "bypass_wgrad_tma") != DENO_OK) return -2;
  prof_before("bypass_wgrad_tma", st);
  DENO_LAUNCH((tc::bypass_wgrad_tma_kernel), dim3(grid), dim3(tc::kWtThreads), smem, st, tm_a, tm_b, p);
  if (check_launch("bypass_wgrad_tma", st) != DENO_OK) return -2;
  return grid;
}

bool wgrad_tc_eligible(const Shape& s) {
  if (path_pref() == PathPref::Simt) return false;
  const int COs = round_up(s.CO, 8), CIs = round_up(s.CI, 8);
  if (COs > 128 || CIs > 128) return false;
  int NS = 4;
  while (NS > 1 && (NS * COs > 128 || NS * CIs > 128)) NS /= 2;
  if ((NS * CIs) % 16 != 0) return false;
  if (tc::wgrad_tc_smem(NS, COs, CIs).total > kMaxSmemBytes - 1024) return false;
  return (NS * (s.CO + s.CI) * (tc::kWgradKT / 4) + tc::kWgradLoaders - 1) / tc::kWgradLoaders <= 8;
}
#else
bool wgrad_tc_eligible(const Shape&) { return false; }
#endif

// `bias_done`: the bias gradient was already produced from the DC mode by the fused mix kernel.
int launch_bypass_wgrad(const Shape& s, const float* gz, const float* x, const float2* ghat, float* partial, float* glw,
                        float* glb, bool bias_done, cudaStream_t st) {
  WgradCfg c;
  int rc = wgrad_cfg(s, c);
  if (rc != DENO_OK) return rc;
#ifndef DENO_EMULATE
  if (path_pref() != PathPref::Simt) {
    int nparts = try_launch_bypass_wgrad_tma(s, gz, x, partial, st);
    if (nparts == -1) nparts = try_launch_bypass_wgrad_tc(s, gz, x, partial, st);
    if (nparts == -2) return DENO_E_CUDA;
    if (nparts > 0) {
      WgradReduceParams r;
      r.partial = partial; r.ncta = nparts;
      r.nw = s.CO * s.CI; r.nj = r.nw; r.pitch = r.nj;   // the tensor-core partials carry no bias column
      r.glw = glw; r.glb = nullptr;
      if (glw != nullptr) {
        prof_before("bypass_wgrad_reduce", st);
        DENO_LAUNCH(bypass_wgrad_reduce_kernel, dim3((r.nj + 31) / 32), dim3(32, kReduceSlices), kReduceSlices * 32 * sizeof(float), st, r);
        rc = check_launch("bypass_wgrad_reduce", st);
        if (rc != DENO_OK) return rc;
      }
      if (glb != nullptr && !bias_done) {
        tc::BiasDcParams bp;
        bp.ghat = ghat; bp.glb = glb; bp.B = s.B; bp.CO = s.CO; bp.M = s.M; bp.s_total = (float)s.S;
        prof_before("bias_from_dc", st);
        DENO_LAUNCH((tc::bias_from_dc_kernel), dim3((s.CO + 127) / 128), dim3(128), 0, st, bp);
        rc = check_launch("bias_from_dc", st);
      }
      return rc;
    }
  }
#endif
  WgradParams p;
  p.gx = make_geom(s, s.CI);
  p.CO = s.CO;
  p.sbo_g = (long long)s.CO * s.S;
  p.gz = gz; p.x = x; p.partial = partial;
  p.PTS = c.PTS; p.tiles = c.tiles;
  rc = allow_smem(bypass_wgrad_partial_kernel, c.smem, "bypass_wgrad");
  if (rc != DENO_OK) return rc;
  prof_before("bypass_wgrad_partial", st);
  DENO_LAUNCH(bypass_wgrad_partial_kernel, dim3(c.ncta), dim3(c.threads), c.smem, st, p);
  rc = check_launch("bypass_wgrad_partial", st);
  if (rc != DENO_OK) return rc;
  WgradReduceParams r;
  r.partial = partial; r.ncta = c.ncta;
  r.nw = s.CO * s.CI; r.nj = r.nw + s.CO; r.pitch = r.nj;
  r.glw = glw; r.glb = glb;
  prof_before("bypass_wgrad_reduce", st);
  DENO_LAUNCH(bypass_wgrad_reduce_kernel, dim3((r.nj + 31) / 32), dim3(32, kReduceSlices), kReduceSlices * 32 * sizeof(float), st, r);
  return check_launch("bypass_wgrad_reduce", st);
}

// ------------------------------------------------------------------------------------------
// Point-wise ends of the stack (include/deno_fno.h, kernels in fno_pointwise.cuh)
// ------------------------------------------------------------------------------------------
struct FnoShape {
  PointGeom g;
  int ndim, B, C, Hd, Kx, ND, O;
  int pdim[3];      // padded extents, outermost first (as in the descriptor)
};

bool proj_width_built(int C) {
  switch (C) {
    case 8: case 12: case 16: case 20: case 24: case 32: case 40: case 48: case 64: return true;
    default: return false;
  }
}

int parse_fno(const deno_fno_desc* d, FnoShape& f) {
  if (d == nullptr) return fail(DENO_E_INVALID, "null descriptor");
  if (d->ndim < 1 || d->ndim > 3) return fail(DENO_E_INVALID, "ndim must be 1, 2 or 3");
  if (d->batch < 1 || d->width < 1 || d->hidden < 1 || d->in_feats < 1 || d->grid_feats < 0 || d->out_dim < 1)
    return fail(DENO_E_INVALID, "batch/width/hidden/in_feats/out_dim must be positive");
  if (d->padding < 0) return fail(DENO_E_INVALID, "padding must be non-negative");
  int sdim[3] = {1, 1, 1};
  long long npts = 1, nppts = 1;
  for (int a = 0; a < d->ndim; ++a) {
    if (d->spatial[a] < 1) return fail(DENO_E_INVALID, "spatial sizes must be positive");
    sdim[3 - d->ndim + a] = d->spatial[a];
    npts *= d->spatial[a];
    nppts *= d->spatial[a] + d->padding;
  }
  if ((long long)d->batch * nppts * (d->hidden > d->width ? d->hidden : d->width) >= (1LL << 40) || nppts >= (1LL << 31))
    return fail(DENO_E_UNSUPPORTED, "grid too large");
  if (d->in_feats + d->grid_feats > kLiftMaxK) return fail(DENO_E_UNSUPPORTED, "more than 16 input features (x ++ grid)");
  if (d->out_dim > kProjMaxO) return fail(DENO_E_UNSUPPORTED, "out_dim above 4");
  if (!proj_width_built(d->width)) return fail(DENO_E_UNSUPPORTED, "width is not one of the built point-wise kernels (8..64, multiple of 4)");
  if (d->hidden > 1024) return fail(DENO_E_UNSUPPORTED, "hidden above 1024");
  f.ndim = d->ndim; f.B = d->batch; f.C = d->width; f.Hd = d->hidden; f.Kx = d->in_feats; f.ND = d->grid_feats; f.O = d->out_dim;
  f.g.sx = sdim[0]; f.g.sy = sdim[1]; f.g.sz = sdim[2];
  f.g.px = d->ndim >= 3 ? sdim[0] + d->padding : 1;
  f.g.py = d->ndim >= 2 ? sdim[1] + d->padding : 1;
  f.g.pz = sdim[2] + d->padding;
  f.g.npts = (int)npts;
  f.g.nppts = (int)nppts;
  for (int a = 0; a < 3; ++a) f.pdim[a] = a < d->ndim ? d->spatial[a] + d->padding : 0;
  return DENO_OK;
}

// The fc1 weight gradient is gW[j, c] = sum_{b, pts} gt[b, j, pt] h[b, c, pt]: the bypass-conv weight gradient of a
// layer with CI = width, CO = hidden on the padded grid.
Shape proj_wgrad_shape(const FnoShape& f) {
  Shape s;
  memset(&s, 0, sizeof(s));
  s.ndim = 1; s.B = f.B; s.CI = f.C; s.CO = f.Hd;
  s.X = 1; s.H = 1; s.W = f.g.nppts;                // one flat "plane": the bypass_wgrad kernels only see points
  s.mX = s.mH = s.mW = 1;
  s.KX = s.KH = s.MW = 1;
  s.KX4 = s.KH4 = s.MWp = 4;
  s.S = f.g.nppts;
  s.Q = 1; s.M = 1; s.act = 0;
  s.simt_only = false;
  return s;
}

struct ProjBwdWs { size_t gt, small, wpart, total; int nblk, nsmall; };
int proj_bwd_ws(const FnoShape& f, ProjBwdWs& w) {
  WgradCfg wc;
  const Shape s = proj_wgrad_shape(f);
  int rc = wgrad_cfg(s, wc);
  if (rc != DENO_OK) return rc;
  const int per = kProjThreads * proj_bwd_pt(f.C);
  w.nblk = f.B * ((f.g.nppts + per - 1) / per);
  w.nsmall = f.Hd + f.O * f.Hd + f.O;
  size_t off = 0;
  w.gt = off;    off += align_up((size_t)f.B * f.Hd * ((f.g.nppts + 127) / 128 * 128) * 4);   // plain or 128-point blocks
  // one partial row per CTA of whichever backward kernel runs: the SIMT kernel launches nblk CTAs, the persistent
  // tensor-core kernel min(tiles, kProjTcCtas) - size for the larger count so neither can run into `wpart`
  const int nsmall_rows = w.nblk > kProjTcCtas ? w.nblk : kProjTcCtas;
  w.small = off; off += align_up((size_t)nsmall_rows * w.nsmall * 4);
  const size_t nparts = (size_t)(wc.ncta > kWgradTcMaxCtas ? wc.ncta : kWgradTcMaxCtas);
  w.wpart = off; off += align_up(nparts * ((size_t)f.Hd * f.C + f.Hd) * 4);
  w.total = off;
  return DENO_OK;
}

struct LiftBwdWs { size_t partial, total; int nblk, nj; };
void lift_bwd_ws(const FnoShape& f, LiftBwdWs& w) {
  w.nblk = f.B * ((f.g.nppts + kLiftThreads - 1) / kLiftThreads);
  w.nj = f.C * (f.Kx + f.ND + 1);
  w.partial = 0;
  w.total = align_up((size_t)w.nblk * w.nj * 4);
}

// Sums `nparts` partial rows (pitch floats apart) of nj columns in a fixed order; the first nw sums go to out_w, the rest to out_b.
int launch_reduce(const float* partial, int nparts, int pitch, int nj, int nw, float* out_w, float* out_b, const char* label,
                  cudaStream_t st) {
  WgradReduceParams r;
  r.partial = partial; r.ncta = nparts; r.pitch = pitch; r.nj = nj; r.nw = nw; r.glw = out_w; r.glb = out_b;
  prof_before(label, st);
  DENO_LAUNCH(bypass_wgrad_reduce_kernel, dim3((nj + 31) / 32), dim3(32, kReduceSlices), kReduceSlices * 32 * sizeof(float), st, r);
  return check_launch(label, st);
}

#ifndef DENO_EMULATE
// Tensor-core projection kernels (fno_proj_tc.cuh): 0 = launched, -1 = shape not covered, else an error code.
tc::ProjTcParams proj_tc_params(const FnoShape& f) {
  tc::ProjTcParams p;
  memset(&p, 0, sizeof(p));
  p.g = f.g; p.B = f.B; p.O = f.O;
  p.tiles_per_sample = (f.g.nppts + 127) / 128;
  p.ntiles = f.B * p.tiles_per_sample;
  return p;
}

template <int C>
int launch_proj_fwd_tc(const tc::ProjTcParams& p, cudaStream_t st) {
  using S = tc::ProjTcSmem<C, 128>;
  int rc = allow_smem(tc::fno_proj_fwd_tc_kernel<C, 128>, (size_t)S::fwd_total, "fno_proj_fwd_tc");
  if (rc != DENO_OK) return rc;
  const int grid = p.ntiles < 148 ? p.ntiles : 148;
  prof_before("fno_proj_fwd_tc", st);
  DENO_LAUNCH((tc::fno_proj_fwd_tc_kernel<C, 128>), dim3(grid), dim3(tc::kPjFwdThreads), (size_t)S::fwd_total, st, p);
  return check_launch("fno_proj_fwd_tc", st);
}

int try_launch_proj_fwd_tc(const FnoShape& f, tc::ProjTcParams p, cudaStream_t st) {
  if (path_pref() == PathPref::Simt || f.Hd != 128 || p.tiles_per_sample > 4096 || f.B >= (1 << 19)) return -1;
  if (f.C == 32) return launch_proj_fwd_tc<32>(p, st);
  if (f.C == 64) return launch_proj_fwd_tc<64>(p, st);
  return -1;
}

// The tensor-core backward writes gt in blocks of 128 points (strided 512-byte runs cost 2.4x in the store path,
// profiles/r2m_phase_probe.txt), which only the tensor-core bypass_wgrad kernel reads: both or neither.
bool proj_bwd_tc_eligible(const FnoShape& f) {
  return path_pref() != PathPref::Simt && f.Hd == 128 && f.C == 32 && f.O == 1 && (f.g.nppts + 127) / 128 <= 4096 &&
         f.B < (1 << 19) && wgrad_tc_eligible(proj_wgrad_shape(f));
}
int launch_proj_bwd_tc(const tc::ProjTcParams& p, int& nparts, cudaStream_t st) {
  using S = tc::ProjTcSmem<32, 128>;
  int rc = allow_smem(tc::fno_proj_bwd_tc_kernel<32, 128>, (size_t)S::bwd_total, "fno_proj_bwd_tc");
  if (rc != DENO_OK) return rc;
  const int grid = p.ntiles < kProjTcCtas ? p.ntiles : kProjTcCtas;
  nparts = grid;
  prof_before("fno_proj_bwd_tc", st);
  DENO_LAUNCH((tc::fno_proj_bwd_tc_kernel<32, 128>), dim3(grid), dim3(tc::kPjThreads), (size_t)S::bwd_total, st, p);
  return check_launch("fno_proj_bwd_tc", st);
}
#endif

#define DENO_PROJ_DISPATCH(C_, CALL) \
  switch (C_) {                      \
    case 8: CALL(8) break;           \
    case 12: CALL(12) break;         \
    case 16: CALL(16) break;         \
    case 20: CALL(20) break;         \
    case 24: CALL(24) break;         \
    case 32: CALL(32) break;         \
    case 40: CALL(40) break;         \
    case 48: CALL(48) break;         \
    case 64: CALL(64) break;         \
    default: return fail(DENO_E_UNSUPPORTED, "width not built"); \
  }

// ------------------------------------------------------------------------------------------
// Adaptive Fourier layers (include/deno_afno.h)
// ------------------------------------------------------------------------------------------
struct AfnoShape {
  Shape s;            // transform geometry: CI = CO = C, every frequency of the leading axes, modes_last bins of the last
  int NB, BS;
  float lambda;
};

// AdaptiveFourier{1,2,3}d (/root/reference/Models/fno/spectral_layers.py:341-573).  The leading axes keep ALL their
// frequencies as a one-sided table f_k = k (KH = H, KX = X: mH = H makes every `k < mH ? k : ...` pick f = k), the last
// axis its first modes_last half-spectrum bins.
int parse_afno(const deno_afno_desc* d, AfnoShape& a) {
  if (d == nullptr) return fail(DENO_E_INVALID, "null descriptor");
  if (d->ndim < 1 || d->ndim > 3) return fail(DENO_E_INVALID, "ndim must be 1, 2 or 3");
  if (d->batch < 1 || d->channels < 1 || d->num_blocks < 1) return fail(DENO_E_INVALID, "batch/channels/num_blocks must be positive");
  if (d->channels % d->num_blocks != 0)
    return fail(DENO_E_INVALID, "hidden_size should be divisible by num_blocks (the reference asserts, spectral_layers.py:352)");
  if (d->flags != 0) return fail(DENO_E_INVALID, "flags must be 0");
  if (d->act < DENO_ACT_IDENTITY || d->act > DENO_ACT_LEAKYRELU) return fail(DENO_E_INVALID, "unknown activation id");
  if (!(d->sparsity_threshold >= 0.0f)) return fail(DENO_E_INVALID, "sparsity_threshold must be >= 0 (softshrink)");
  for (int k = 0; k < d->ndim; ++k)
    if (d->spatial[k] < 1) return fail(DENO_E_INVALID, "spatial sizes must be positive");
  const int nd = d->ndim;
  Shape& s = a.s;
  memset(&s, 0, sizeof(s));
  s.ndim = nd;
  s.B = d->batch; s.CI = s.CO = d->channels; s.act = d->act;
  s.W = d->spatial[nd - 1];
  s.mW = d->modes_last;
  if (s.mW < 1 || s.mW > s.W / 2 + 1) return fail(DENO_E_INVALID, "modes_last must be 1 .. spatial[last] / 2 + 1");
  s.H = nd >= 2 ? d->spatial[nd - 2] : 1;
  s.X = nd == 3 ? d->spatial[0] : 1;
  s.mH = nd >= 2 ? s.H : 0;
  s.mX = nd == 3 ? s.X : 0;
  s.KH = nd >= 2 ? s.H : 1;
  s.KX = nd == 3 ? s.X : 1;
  s.MW = s.mW;
  s.KH4 = round_up(s.KH, 4);
  s.KX4 = round_up(s.KX, 4);
  s.MWp = round_up(s.MW, 4);
  s.S = (long long)s.X * s.H * s.W;
  s.Q = s.KH * s.MW;
  if ((long long)s.KX * s.Q > 0x7fffffffLL) return fail(DENO_E_UNSUPPORTED, "spectrum too large");
  s.M = s.KX * s.Q;
  s.simt_only = true;
  a.NB = d->num_blocks;
  a.BS = d->channels / d->num_blocks;
  a.lambda = d->sparsity_threshold;
  if (a.BS > kAfnoMaxBlock) return fail(DENO_E_UNSUPPORTED, "block size (hidden_size / num_blocks) above 64 is not built");
  return DENO_OK;
}

// Launch geometry of the block MLP kernels (pure function of the shape: workspace sizing must agree).
struct AfnoCfg { int TP, tpb, chunks; size_t smem; };
int afno_cfg(const AfnoShape& a, bool backward, AfnoCfg& c) {
  const int threads = backward ? kAfnoBwdThreads : kAfnoFwdThreads;
  const int ntile_bufs = backward ? 4 : 2;
  int TP = threads;
  while (TP > 32 && afno_smem_bytes(a.BS, TP, ntile_bufs) > (size_t)200 * 1024) TP /= 2;
  while (TP > 32 && TP / 2 >= a.s.M) TP /= 2;                    // tiny spectra: no point in mostly-empty tiles
  c.TP = TP;
  c.smem = afno_smem_bytes(a.BS, TP, ntile_bufs);
  if (c.smem > (size_t)kMaxSmemBytes) return fail(DENO_E_UNSUPPORTED, "afno block MLP does not fit shared memory");
  c.tpb = (a.s.M + TP - 1) / TP;
  const long long ntiles = (long long)a.s.B * c.tpb;
  long long chunks = (148LL * 4 + a.NB - 1) / a.NB;              // about four CTAs per SM over all blocks
  if (chunks > ntiles) chunks = ntiles;
  if (chunks < 1) chunks = 1;
  c.chunks = (int)chunks;
  return DENO_OK;
}

struct AfnoWs { size_t xhat, yhat, t2, v, partial, total; };
int afno_ws(const AfnoShape& a, bool backward, AfnoWs& w) {
  const Shape& s = a.s;
  AfnoCfg c;
  int rc = afno_cfg(a, backward, c);
  if (rc != DENO_OK) return rc;
  const size_t spec = align_up((size_t)s.B * s.CI * s.M * 8);
  size_t off = 0;
  w.xhat = off; off += spec;                                     // forward: X^ (unless saved by the caller); backward: G^
  w.yhat = off; off += spec;                                     // forward: filtered spectrum; backward: G^x
  w.t2 = off;   off += (s.ndim == 3) ? align_up((size_t)s.B * s.X * s.CI * s.Q * 8) : 0;
  w.v = off;    off += (s.ndim == 3) ? align_up((size_t)s.B * s.X * s.CI * s.Q * 8) : 0;
  w.partial = off; off += backward ? align_up((size_t)c.chunks * afno_partial(a.NB, a.BS).total * 4) : 0;
  w.total = off;
  return DENO_OK;
}

// identity bypass of the residual `+ x`: a C x C unit matrix behind the DFT tables of the transform shape
size_t afno_eye_offset(const Shape& s) { return (tw_layout(s).total + 63) / 64 * 64; }

int launch_afno_mlp(const AfnoShape& a, bool backward, const float2* xin, const float2* gin, float2* out, const void* w1,
                    const void* b1, const void* w2, const void* b2, float* partial, cudaStream_t st) {
  AfnoCfg c;
  int rc = afno_cfg(a, backward, c);
  if (rc != DENO_OK) return rc;
  AfnoParams p;
  p.xin = xin; p.gin = gin; p.out = out;
  p.w1 = static_cast<const float2*>(w1); p.b1 = static_cast<const float2*>(b1);
  p.w2 = static_cast<const float2*>(w2); p.b2 = static_cast<const float2*>(b2);
  p.partial = partial;
  p.B = a.s.B; p.C = a.s.CI; p.NB = a.NB; p.BS = a.BS; p.M = a.s.M;
  p.act = a.s.act; p.lambda = a.lambda;
  p.TP = c.TP; p.tpb = c.tpb; p.pitch = afno_partial(a.NB, a.BS).total;
  const dim3 grid((unsigned)c.chunks, (unsigned)a.NB);
  if (!backward) {
    rc = allow_smem(afno_mlp_fwd_kernel, c.smem, "afno_mlp_fwd");
    if (rc != DENO_OK) return rc;
    prof_before("afno_mlp_fwd", st);
    DENO_LAUNCH(afno_mlp_fwd_kernel, grid, dim3(kAfnoFwdThreads), c.smem, st, p);
    return check_launch("afno_mlp_fwd", st);
  }
  rc = allow_smem(afno_mlp_bwd_kernel, c.smem, "afno_mlp_bwd");
  if (rc != DENO_OK) return rc;
  prof_before("afno_mlp_bwd", st);
  DENO_LAUNCH(afno_mlp_bwd_kernel, grid, dim3(kAfnoBwdThreads), c.smem, st, p);
  return check_launch("afno_mlp_bwd", st);
}

}  // namespace

// ==========================================================================================
// extern "C" surface
// ==========================================================================================
extern "C" {

int deno_spectral_abi_version(void) { return DENO_SPECTRAL_ABI_VERSION; }

int deno_reload_env(void) {
  std::lock_guard<std::mutex> lock(g_opt_mutex);
  g_options = read_options();
  g_options_read = true;
  return DENO_OK;
}

const char* deno_last_error(void) { return g_error.c_str(); }

int deno_spectral_last_launch_count(void) { return g_launches; }

int deno_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, const float* step, void* stream) {
  g_launches = 0;
  if (!p || !g || !m || !v || !step) return fail(DENO_E_INVALID, "deno_adam_step: null pointer argument");
  if (n == 0) return DENO_OK;
  AdamParams a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.n = (long long)n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long blocks = ((long long)n + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  prof_before("adam_flat", st);
  DENO_LAUNCH(adam_flat_kernel, dim3((unsigned)blocks), dim3(256), 0, st, a);
  return check_launch("adam_flat", st);
}

int deno_dp_fused_adam(float* p_local, float* p_mc, float* g_mc, float* m, float* v, size_t n, int rank, int world,
                       void* signal_pads, int write_back_grads, float lr, float beta1, float beta2, float eps,
                       float weight_decay, const float* step, void* stream) {
  g_launches = 0;
  if (!p_local || !p_mc || !g_mc || !m || !v || !signal_pads || !step)
    return fail(DENO_E_INVALID, "deno_dp_fused_adam: null pointer argument");
  if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(DENO_E_INVALID, "deno_dp_fused_adam: bad rank / world");
  if ((n % 4) != 0) return fail(DENO_E_INVALID, "deno_dp_fused_adam: n must be a multiple of 4 floats");
  if (((reinterpret_cast<uintptr_t>(p_local) | reinterpret_cast<uintptr_t>(p_mc) | reinterpret_cast<uintptr_t>(g_mc) |
        reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) != 0)
    return fail(DENO_E_INVALID, "deno_dp_fused_adam: buffers must be 16-byte aligned");
  if (n == 0) return DENO_OK;
#ifndef DENO_EMULATE
  DpFusedParams a;
  a.p_local = p_local; a.p_mc = p_mc; a.g_mc = g_mc; a.m = m; a.v = v; a.n = (long long)n;
  a.rank = rank; a.world = world; a.signal_pads = static_cast<uint32_t**>(signal_pads);
  a.write_back_grads = write_back_grads;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.step = step;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long per4 = ((long long)n / 4 + world - 1) / world;          // float4 of one rank's slice
  long long blocks = (per4 + (long long)kDpInFlight * kDpThreads - 1) / ((long long)kDpInFlight * kDpThreads);
  if (blocks > kDpMaxBlocks) blocks = kDpMaxBlocks;
  if (blocks * world > 512) blocks = 512 / world;                          // slots of a 2 KB signal pad
  if (blocks < 1) blocks = 1;
  prof_before("dp_allreduce_adam", st);
  // the blocks of all ranks wait for each other: never more blocks than can be resident (32 << 148 SMs)
  dp_allreduce_adam_kernel<<<dim3((unsigned)blocks), dim3(kDpThreads), 0, st>>>(a);
  return check_launch("dp_allreduce_adam", st);
#else
  (void)lr; (void)beta1; (void)beta2; (void)eps; (void)weight_decay; (void)write_back_grads; (void)stream;
  return fail(DENO_E_UNSUPPORTED, "deno_dp_fused_adam: not available in the CPU emulation build");
#endif
}

int deno_debug_read(long long* host, int count, int clear) {
#if !defined(DENO_EMULATE) && defined(DENO_PHASE_PROBE)
  if (host == nullptr || count < 0 || count > 64) return fail(DENO_E_INVALID, "deno_debug_read: bad arguments");
  cudaError_t e = cudaMemcpyFromSymbol(host, tc::g_dbg, sizeof(long long) * count);
  if (e != cudaSuccess) return fail(DENO_E_CUDA, cudaGetErrorString(e));
  if (clear) {
    long long zeros[64] = {0};
    e = cudaMemcpyToSymbol(tc::g_dbg, zeros, sizeof(zeros));
    if (e != cudaSuccess) return fail(DENO_E_CUDA, cudaGetErrorString(e));
  }
#else
  if (host == nullptr || count < 0 || count > 64) return fail(DENO_E_INVALID, "deno_debug_read: bad arguments");
  for (int i = 0; i < count; ++i) host[i] = 0;   // product build: no phase probe compiled in
  (void)clear;
#endif
  return DENO_OK;
}

int deno_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_prof.clear();
  g_profiling = true;
  return DENO_OK;
}

int deno_profile_end(deno_profile_record* records, int capacity) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_profiling = false;
  const int n = (int)g_prof.size();
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
#ifndef DENO_EMULATE
    cudaEventSynchronize(g_prof[i].stop);
    cudaEventElapsedTime(&ms, g_prof[i].start, g_prof[i].stop);
    cudaEventDestroy(g_prof[i].start);
    cudaEventDestroy(g_prof[i].stop);
#endif
    if (records != nullptr && i < capacity) {
      strncpy(records[i].name, g_prof[i].name, sizeof(records[i].name) - 1);
      records[i].name[sizeof(records[i].name) - 1] = 0;
      records[i].ms = ms;
    }
  }
  g_prof.clear();
  return n;
}

size_t deno_spectral_num_modes(const deno_spectral_desc* desc) {
  Shape s;
  if (parse_desc(desc, s) != DENO_OK) return 0;
  return (size_t)s.M;
}

size_t deno_spectral_twiddle_bytes(const deno_spectral_desc* desc) {
  Shape s;
  if (parse_desc(desc, s) != DENO_OK) return 0;
  return tw_layout(s).total * sizeof(float);
}

static int twiddle_fill_shape(const Shape& s, void* host_buf, size_t bytes) {
  const TwLayout t = tw_layout(s);
  if (host_buf == nullptr || bytes < t.total * sizeof(float)) return fail(DENO_E_INVALID, "twiddle buffer too small");
  float* base = static_cast<float*>(host_buf);
  fill_axis(base + t.last, s.W, s.MW, s.MWp, s.mW, /*half=*/true);
  if (s.ndim >= 2) {
    fill_axis(base + t.mid, s.H, s.KH, s.KH4, s.mH, /*half=*/s.ndim == 4);   // 4-D: the third axis keeps its low block only
  } else {
    fill_axis(base + t.mid, 1, 1, s.KH4, 0, true);
  }
  if (s.ndim == 3) fill_axis(base + t.outer, s.X, s.KX, s.KX4, s.mX, false);
  if (s.ndim == 4) {
    fill_axis(base + t.outer, s.X1, s.KX1, s.KX14, s.mX1, false);
    fill_axis(base + t.outer2, s.X2, s.KX2, s.KX24, s.mX2, false);
  }
  fill_synth_ae(base + t.synth_ae, s, t);
  fill_analysis_images(base, s, t);
  memset(base + t.at_e1, 0, (t.total - t.at_e1) * sizeof(float));
  fill_analysis_tma_images(base, s, t);
  fill_row_synth_images(base, s, t);
  for (int l = 0; l < s.MWp; ++l) {
    float c = 0.f;
    if (l < s.MW) {
      c = 2.f;
      if (l == 0) c = 1.f;
      if ((s.W % 2) == 0 && l == s.W / 2) c = 1.f;
    }
    base[t.cl + l] = c;
  }
  return DENO_OK;
}

int deno_spectral_twiddle_fill(const deno_spectral_desc* desc, void* host_buf, size_t bytes) {
  Shape s;
  int rc = parse_desc(desc, s);
  if (rc != DENO_OK) return rc;
  return twiddle_fill_shape(s, host_buf, bytes);
}

// ---- 4-D stand-alone layer: the same queries on a deno_spectral_desc4 ----
size_t deno_spectral4_num_modes(const deno_spectral_desc4* desc) {
  Shape s;
  if (parse_desc4(desc, s) != DENO_OK) return 0;
  return (size_t)s.M;
}
size_t deno_spectral4_twiddle_bytes(const deno_spectral_desc4* desc) {
  Shape s;
  if (parse_desc4(desc, s) != DENO_OK) return 0;
  return tw_layout(s).total * sizeof(float);
}
int deno_spectral4_twiddle_fill(const deno_spectral_desc4* desc, void* host_buf, size_t bytes) {
  Shape s;
  int rc = parse_desc4(desc, s);
  if (rc != DENO_OK) return rc;
  return twiddle_fill_shape(s, host_buf, bytes);
}
size_t deno_spectral4_workspace_bytes(const deno_spectral_desc4* desc, int backward) {
  Shape s;
  if (parse_desc4(desc, s) != DENO_OK) return 0;
  if (!backward) return fwd_ws(s).total;
  BwdWs w;
  if (bwd_ws(s, w) != DENO_OK) return 0;
  return w.total;
}

size_t deno_spectral_workspace_bytes(const deno_spectral_desc* desc, int backward) {
  Shape s;
  if (parse_desc(desc, s) != DENO_OK) return 0;
  if (!backward) return fwd_ws(s).total;
  BwdWs w;
  if (bwd_ws(s, w) != DENO_OK) return 0;
  return w.total;
}

size_t deno_spectral_xhat_bytes(const deno_spectral_desc* desc) {
  Shape s;
  if (parse_desc(desc, s) != DENO_OK) return 0;
  return (size_t)s.B * s.CI * s.M * 8;
}

static int spectral_fwd_shape(const Shape& s, const float* x, const void* const* w_blocks,
                              const float* lin_w, const float* lin_b, const void* twiddles, float* y,
                              float* z_saved, void* xhat_saved, void* workspace, size_t workspace_bytes,
                              void* stream) {
  int rc;
  if (!x || !w_blocks || !lin_w || !twiddles || !y || !workspace) return fail(DENO_E_INVALID, "null pointer argument");
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < nblk; ++j)
    if (w_blocks[j] == nullptr) return fail(DENO_E_INVALID, "null weight block");
  const FwdWs ws = fwd_ws(s);
  if (workspace_bytes < ws.total) return fail(DENO_E_WORKSPACE, "workspace too small for deno_spectral_fwd");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* wsb = static_cast<char*>(workspace);
  const TwPtrs tw = tw_ptrs(s, twiddles);
  float2* xhat = xhat_saved ? static_cast<float2*>(xhat_saved) : reinterpret_cast<float2*>(wsb + ws.xhat);
  float2* yhat = reinterpret_cast<float2*>(wsb + ws.yhat);

  if (s.ndim >= 3) {
    float2* t2 = reinterpret_cast<float2*>(wsb + ws.t2);
    float2* v = reinterpret_cast<float2*>(wsb + ws.v);
    if ((rc = launch_analysis(s, s.CI, x, nullptr, nullptr, t2, tw, 1.0f, 0, st))) return rc;
    if (s.ndim == 4) {
      if ((rc = launch_outer4(s, s.CI, true, t2, reinterpret_cast<float2*>(wsb + ws.t3), xhat, tw, st))) return rc;
    } else if ((rc = launch_outer(s, s.CI, true, t2, xhat, tw, st))) return rc;
    if ((rc = launch_mix(s, false, xhat, yhat, w_blocks, st))) return rc;
    if (s.ndim == 4) {
      if ((rc = launch_outer4(s, s.CO, false, yhat, reinterpret_cast<float2*>(wsb + ws.v3), v, tw, st))) return rc;
    } else if ((rc = launch_outer(s, s.CO, false, yhat, v, tw, st))) return rc;
    return launch_synthesis(s, s.CI, s.CO, v, x, lin_w, s.CI, 1, lin_b, y, z_saved, tw, (float)(1.0 / (double)s.S), 1,
                            s.act, reinterpret_cast<float*>(wsb + ws.urow), st);
  }
  if ((rc = launch_analysis(s, s.CI, x, nullptr, nullptr, xhat, tw, 1.0f, 0, st))) return rc;
  if ((rc = launch_mix(s, false, xhat, yhat, w_blocks, st))) return rc;
  return launch_synthesis(s, s.CI, s.CO, yhat, x, lin_w, s.CI, 1, lin_b, y, z_saved, tw, (float)(1.0 / (double)s.S), 1,
                          s.act, reinterpret_cast<float*>(wsb + ws.urow), st);
}

int deno_spectral_fwd(const deno_spectral_desc* desc, const float* x, const void* const* w_blocks,
                      const float* lin_w, const float* lin_b, const void* twiddles, float* y,
                      float* z_saved, void* xhat_saved, void* workspace, size_t workspace_bytes,
                      void* stream) {
  g_launches = 0;
  Shape s;
  int rc = parse_desc(desc, s);
  if (rc != DENO_OK) return rc;
  return spectral_fwd_shape(s, x, w_blocks, lin_w, lin_b, twiddles, y, z_saved, xhat_saved, workspace, workspace_bytes, stream);
}

int deno_spectral4_fwd(const deno_spectral_desc4* desc, const float* x, const void* const* w_blocks,
                       const float* lin_w, const float* lin_b, const void* twiddles, float* y,
                       float* z_saved, void* xhat_saved, void* workspace, size_t workspace_bytes,
                       void* stream) {
  g_launches = 0;
  Shape s;
  int rc = parse_desc4(desc, s);
  if (rc != DENO_OK) return rc;
  return spectral_fwd_shape(s, x, w_blocks, lin_w, lin_b, twiddles, y, z_saved, xhat_saved, workspace, workspace_bytes, stream);
}

// Shared body of deno_spectral_bwd / deno_spectral_bwd_overlap.  With `ov` the kernels that only PRODUCE parameter
// gradients (gW of the mode mix, the bypass weight / bias gradient and its reduction) go to ov->side_stream right after
// gz and G^ exist, and the input-gradient chain (G^x -> row synthesis -> plane synthesis) continues on `stream`; the
// two meet again before the call returns.  Both sides only read gz / G^ / X^ / x and write disjoint outputs and
// disjoint workspace regions (BwdWs: gxhat, v, urow on the chain; partial on the side).
static int spectral_bwd_impl(const Shape& s, const float* gy, const float* x,
                             const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                             const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                             float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                             void* stream, const deno_overlap* ov) {
  int rc;
  if (!gy || !x || !z_saved || !xhat_saved || !w_blocks || !lin_w || !twiddles || !workspace)
    return fail(DENO_E_INVALID, "null pointer argument");
  const int nblk = num_weight_blocks(s);
  for (int j = 0; j < nblk; ++j) {
    if (w_blocks[j] == nullptr) return fail(DENO_E_INVALID, "null weight block");
    if (gw_blocks != nullptr && gw_blocks[j] == nullptr) return fail(DENO_E_INVALID, "null weight-gradient block");
  }
  BwdWs ws;
  if ((rc = bwd_ws(s, ws))) return rc;
  if (workspace_bytes < ws.total) return fail(DENO_E_WORKSPACE, "workspace too small for deno_spectral_bwd");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* wsb = static_cast<char*>(workspace);
  const TwPtrs tw = tw_ptrs(s, twiddles);
  float* gz = reinterpret_cast<float*>(wsb + ws.gz);
  float2* ghat = reinterpret_cast<float2*>(wsb + ws.ghat);
  float2* gxhat = reinterpret_cast<float2*>(wsb + ws.gxhat);
  const float2* xhat = static_cast<const float2*>(xhat_saved);
  const float inv_s = (float)(1.0 / (double)s.S);

  // gz = gy * act'(z);  G^ = (c_l / S) * truncated DFT(gz)
  if (s.ndim >= 3) {
    float2* t2 = reinterpret_cast<float2*>(wsb + ws.t2);
    if ((rc = launch_analysis(s, s.CO, gy, z_saved, gz, t2, tw, inv_s, 1, st))) return rc;
    if (s.ndim == 4) {
      if ((rc = launch_outer4(s, s.CO, true, t2, reinterpret_cast<float2*>(wsb + ws.t3), ghat, tw, st))) return rc;
    } else if ((rc = launch_outer(s, s.CO, true, t2, ghat, tw, st))) return rc;
  } else {
    if ((rc = launch_analysis(s, s.CO, gy, z_saved, gz, ghat, tw, inv_s, 1, st))) return rc;
  }
  const bool bias_from_dc = (glin_b != nullptr) && wgrad_tc_eligible(s);
  const bool want_side = gw_blocks != nullptr || glin_w != nullptr || glin_b != nullptr;
  cudaStream_t sd = st;
#ifndef DENO_EMULATE
  // per-launch event bracketing (deno_profile_begin) keeps everything on one stream: same launches, serial
  const bool fork = ov != nullptr && ov->side_stream != nullptr && gx != nullptr && want_side && !g_profiling;
  if (fork) {
    sd = static_cast<cudaStream_t>(ov->side_stream);
    if (cudaEventRecord(static_cast<cudaEvent_t>(ov->ev_fork), st) != cudaSuccess ||
        cudaStreamWaitEvent(sd, static_cast<cudaEvent_t>(ov->ev_fork), 0) != cudaSuccess)
      return fail(DENO_E_CUDA, std::string("bwd overlap fork: ") + cudaGetErrorString(cudaGetLastError()));
  }
#else
  const bool fork = false;
  (void)ov;
#endif
  const bool split = ov != nullptr && gx != nullptr && want_side;   // two role launches instead of the fused one
  if (!split) {
    // one launch: G^x = conj(W) G^ (if gx is wanted), gW (if wanted) and, on the tensor-core wgrad path, the bias gradient
    if ((rc = launch_mix_bwd_fused(s, xhat, ghat, gx != nullptr ? gxhat : nullptr, w_blocks, gw_blocks,
                                   bias_from_dc ? glin_b : nullptr, st)))
      return rc;
  } else {
    if ((rc = launch_mix_bwd_fused(s, xhat, ghat, gxhat, w_blocks, nullptr, nullptr, st, "mode_mix_bwd_gx"))) return rc;
  }
  if (gx != nullptr) {
    const float2* spec = gxhat;
    if (s.ndim >= 3) {
      float2* v = reinterpret_cast<float2*>(wsb + ws.v);
      if (s.ndim == 4) {
        if ((rc = launch_outer4(s, s.CI, false, gxhat, reinterpret_cast<float2*>(wsb + ws.v3), v, tw, st))) return rc;
      } else if ((rc = launch_outer(s, s.CI, false, gxhat, v, tw, st))) return rc;
      spec = v;
    }
    // gx = Re sum G^x e^{+i theta} + W_l^T gz   (channels swap roles: "input" = gz with CO channels)
    if ((rc = launch_synthesis(s, s.CO, s.CI, spec, gz, lin_w, 1, s.CI, nullptr, gx, nullptr, tw, 1.0f, 0,
                               DENO_ACT_IDENTITY, reinterpret_cast<float*>(wsb + ws.urow), st)))
      return rc;
  }
  if (split && (gw_blocks != nullptr || bias_from_dc)) {
    if ((rc = launch_mix_bwd_fused(s, xhat, ghat, nullptr, nullptr, gw_blocks, bias_from_dc ? glin_b : nullptr, sd,
                                   "mode_mix_bwd_gw")))
      return rc;
  }
  if (glin_w != nullptr || glin_b != nullptr) {
    float* partial = reinterpret_cast<float*>(wsb + ws.partial);
    if ((rc = launch_bypass_wgrad(s, gz, x, ghat, partial, glin_w, glin_b, bias_from_dc, sd))) return rc;
  }
#ifndef DENO_EMULATE
  if (fork) {
    if (cudaEventRecord(static_cast<cudaEvent_t>(ov->ev_join), sd) != cudaSuccess ||
        cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(ov->ev_join), 0) != cudaSuccess)
      return fail(DENO_E_CUDA, std::string("bwd overlap join: ") + cudaGetErrorString(cudaGetLastError()));
  }
#endif
  return DENO_OK;
}

int deno_spectral_bwd(const deno_spectral_desc* desc, const float* gy, const float* x,
                      const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                      const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                      float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                      void* stream) {
  g_launches = 0;
  Shape s;
  int rc = parse_desc(desc, s);
  if (rc != DENO_OK) return rc;
  return spectral_bwd_impl(s, gy, x, z_saved, xhat_saved, w_blocks, lin_w, twiddles, gx, gw_blocks, glin_w, glin_b,
                           workspace, workspace_bytes, stream, nullptr);
}

int deno_spectral_bwd_overlap(const deno_spectral_desc* desc, const float* gy, const float* x,
                              const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                              const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                              float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                              void* stream, const deno_overlap* ov) {
  g_launches = 0;
  Shape s;
  int rc = parse_desc(desc, s);
  if (rc != DENO_OK) return rc;
  return spectral_bwd_impl(s, gy, x, z_saved, xhat_saved, w_blocks, lin_w, twiddles, gx, gw_blocks, glin_w, glin_b,
                           workspace, workspace_bytes, stream, ov);
}

int deno_spectral4_bwd(const deno_spectral_desc4* desc, const float* gy, const float* x,
                       const float* z_saved, const void* xhat_saved, const void* const* w_blocks,
                       const float* lin_w, const void* twiddles, float* gx, void* const* gw_blocks,
                       float* glin_w, float* glin_b, void* workspace, size_t workspace_bytes,
                       void* stream, const deno_overlap* ov) {
  g_launches = 0;
  Shape s;
  int rc = parse_desc4(desc, s);
  if (rc != DENO_OK) return rc;
  return spectral_bwd_impl(s, gy, x, z_saved, xhat_saved, w_blocks, lin_w, twiddles, gx, gw_blocks, glin_w, glin_b,
                           workspace, workspace_bytes, stream, ov);
}

// ---- adaptive Fourier layers (include/deno_afno.h) ----
size_t deno_afno_num_modes(const deno_afno_desc* desc) {
  AfnoShape a;
  if (parse_afno(desc, a) != DENO_OK) return 0;
  return (size_t)a.s.M;
}

size_t deno_afno_twiddle_bytes(const deno_afno_desc* desc) {
  AfnoShape a;
  if (parse_afno(desc, a) != DENO_OK) return 0;
  return (afno_eye_offset(a.s) + (size_t)a.s.CI * a.s.CI) * sizeof(float);
}

int deno_afno_twiddle_fill(const deno_afno_desc* desc, void* host_buf, size_t bytes) {
  AfnoShape a;
  int rc = parse_afno(desc, a);
  if (rc != DENO_OK) return rc;
  const size_t eye = afno_eye_offset(a.s), C = (size_t)a.s.CI;
  if (host_buf == nullptr || bytes < (eye + C * C) * sizeof(float)) return fail(DENO_E_INVALID, "twiddle buffer too small");
  if ((rc = twiddle_fill_shape(a.s, host_buf, bytes))) return rc;
  float* base = static_cast<float*>(host_buf);
  for (size_t k = tw_layout(a.s).total; k < eye; ++k) base[k] = 0.f;
  for (size_t o = 0; o < C; ++o)
    for (size_t i = 0; i < C; ++i) base[eye + o * C + i] = (o == i) ? 1.f : 0.f;
  return DENO_OK;
}

size_t deno_afno_workspace_bytes(const deno_afno_desc* desc, int backward) {
  AfnoShape a;
  if (parse_afno(desc, a) != DENO_OK) return 0;
  AfnoWs w;
  if (afno_ws(a, backward != 0, w) != DENO_OK) return 0;
  return w.total;
}

int deno_afno_fwd(const deno_afno_desc* desc, const float* x, const void* w1, const void* b1, const void* w2,
                  const void* b2, const void* twiddles, float* y, void* xhat_saved, void* workspace,
                  size_t workspace_bytes, void* stream) {
  g_launches = 0;
  AfnoShape a;
  int rc = parse_afno(desc, a);
  if (rc != DENO_OK) return rc;
  if (!x || !w1 || !b1 || !w2 || !b2 || !twiddles || !y || !workspace) return fail(DENO_E_INVALID, "deno_afno_fwd: null pointer argument");
  AfnoWs ws;
  if ((rc = afno_ws(a, false, ws))) return rc;
  if (workspace_bytes < ws.total) return fail(DENO_E_WORKSPACE, "workspace too small for deno_afno_fwd");
  const Shape& s = a.s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* wsb = static_cast<char*>(workspace);
  const TwPtrs tw = tw_ptrs(s, twiddles);
  const float* eye = static_cast<const float*>(twiddles) + afno_eye_offset(s);
  float2* xhat = xhat_saved ? static_cast<float2*>(xhat_saved) : reinterpret_cast<float2*>(wsb + ws.xhat);
  float2* yhat = reinterpret_cast<float2*>(wsb + ws.yhat);
  const float ortho = (float)(1.0 / sqrt((double)s.S));
  // X^ = rfftn(x, norm="ortho"), every frequency of the leading axes
  if (s.ndim == 3) {
    float2* t2 = reinterpret_cast<float2*>(wsb + ws.t2);
    if ((rc = launch_analysis(s, s.CI, x, nullptr, nullptr, t2, tw, ortho, 0, st))) return rc;
    if ((rc = launch_outer(s, s.CI, true, t2, xhat, tw, st))) return rc;
  } else if ((rc = launch_analysis(s, s.CI, x, nullptr, nullptr, xhat, tw, ortho, 0, st))) {
    return rc;
  }
  if ((rc = launch_afno_mlp(a, false, xhat, nullptr, yhat, w1, b1, w2, b2, nullptr, st))) return rc;
  // y = irfftn(., norm="ortho") + x : the residual is the synthesis kernel's bypass with a unit matrix
  const float2* spec = yhat;
  if (s.ndim == 3) {
    float2* v = reinterpret_cast<float2*>(wsb + ws.v);
    if ((rc = launch_outer(s, s.CO, false, yhat, v, tw, st))) return rc;
    spec = v;
  }
  return launch_synthesis(s, s.CI, s.CO, spec, x, eye, s.CI, 1, nullptr, y, nullptr, tw, ortho, 1, DENO_ACT_IDENTITY,
                          nullptr, st);
}

int deno_afno_bwd(const deno_afno_desc* desc, const float* gy, const void* xhat_saved, const void* w1,
                  const void* b1, const void* w2, const void* b2, const void* twiddles, float* gx, void* gw1,
                  void* gb1, void* gw2, void* gb2, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  AfnoShape a;
  int rc = parse_afno(desc, a);
  if (rc != DENO_OK) return rc;
  if (!gy || !xhat_saved || !w1 || !b1 || !w2 || !b2 || !twiddles || !workspace)
    return fail(DENO_E_INVALID, "deno_afno_bwd: null pointer argument");
  const int nw = (gw1 != nullptr) + (gb1 != nullptr) + (gw2 != nullptr) + (gb2 != nullptr);
  if (nw != 0 && nw != 4) return fail(DENO_E_INVALID, "deno_afno_bwd: pass all four parameter gradients or none");
  if (nw == 0 && gx == nullptr) return DENO_OK;
  AfnoWs ws;
  if ((rc = afno_ws(a, true, ws))) return rc;
  if (workspace_bytes < ws.total) return fail(DENO_E_WORKSPACE, "workspace too small for deno_afno_bwd");
  const Shape& s = a.s;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* wsb = static_cast<char*>(workspace);
  const TwPtrs tw = tw_ptrs(s, twiddles);
  const float* eye = static_cast<const float*>(twiddles) + afno_eye_offset(s);
  float2* ghat = reinterpret_cast<float2*>(wsb + ws.xhat);
  float2* gxhat = reinterpret_cast<float2*>(wsb + ws.yhat);
  float* partial = reinterpret_cast<float*>(wsb + ws.partial);
  const float ortho = (float)(1.0 / sqrt((double)s.S));
  // G^ = (c_l / sqrt(S)) DFT(gy): the adjoint of the ortho C2R synthesis
  if (s.ndim == 3) {
    float2* t2 = reinterpret_cast<float2*>(wsb + ws.t2);
    if ((rc = launch_analysis(s, s.CO, gy, nullptr, nullptr, t2, tw, ortho, 1, st))) return rc;
    if ((rc = launch_outer(s, s.CO, true, t2, ghat, tw, st))) return rc;
  } else if ((rc = launch_analysis(s, s.CO, gy, nullptr, nullptr, ghat, tw, ortho, 1, st))) {
    return rc;
  }
  if ((rc = launch_afno_mlp(a, true, static_cast<const float2*>(xhat_saved), ghat, gx != nullptr ? gxhat : nullptr, w1, b1,
                            w2, b2, nw ? partial : nullptr, st)))
    return rc;
  if (nw) {
    AfnoCfg c;
    if ((rc = afno_cfg(a, true, c))) return rc;
    const AfnoPartial pl = afno_partial(a.NB, a.BS);
    const int nwf = a.NB * a.BS * a.BS * 2, nbf = a.NB * a.BS * 2;
    if ((rc = launch_reduce(partial + pl.w1, c.chunks, pl.total, nwf + nbf, nwf, static_cast<float*>(gw1),
                            static_cast<float*>(gb1), "afno_wgrad_reduce", st)))
      return rc;
    if ((rc = launch_reduce(partial + pl.w2, c.chunks, pl.total, nwf + nbf, nwf, static_cast<float*>(gw2),
                            static_cast<float*>(gb2), "afno_wgrad_reduce", st)))
      return rc;
  }
  if (gx != nullptr) {
    // gx = (1 / sqrt(S)) Re sum G^x e^{+i theta} + gy   (no c_l: adjoint of the R2C analysis; the residual's identity)
    const float2* spec = gxhat;
    if (s.ndim == 3) {
      float2* v = reinterpret_cast<float2*>(wsb + ws.v);
      if ((rc = launch_outer(s, s.CI, false, gxhat, v, tw, st))) return rc;
      spec = v;
    }
    if ((rc = launch_synthesis(s, s.CO, s.CI, spec, gy, eye, s.CI, 1, nullptr, gx, nullptr, tw, ortho, 0, DENO_ACT_IDENTITY,
                               nullptr, st)))
      return rc;
  }
  return DENO_OK;
}

int deno_overlap_create(deno_overlap* out) {
  if (out == nullptr) return fail(DENO_E_INVALID, "null deno_overlap");
  memset(out, 0, sizeof(*out));
#ifndef DENO_EMULATE
  cudaStream_t sd = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e0, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess) {
    const std::string msg = std::string("deno_overlap_create: ") + cudaGetErrorString(cudaGetLastError());
    if (e1) cudaEventDestroy(e1);
    if (e0) cudaEventDestroy(e0);
    if (sd) cudaStreamDestroy(sd);
    return fail(DENO_E_CUDA, msg);
  }
  out->side_stream = sd; out->ev_fork = e0; out->ev_join = e1;
#endif
  return DENO_OK;
}

int deno_overlap_destroy(deno_overlap* ov) {
  if (ov == nullptr) return DENO_OK;
#ifndef DENO_EMULATE
  if (ov->ev_join) cudaEventDestroy(static_cast<cudaEvent_t>(ov->ev_join));
  if (ov->ev_fork) cudaEventDestroy(static_cast<cudaEvent_t>(ov->ev_fork));
  if (ov->side_stream) cudaStreamDestroy(static_cast<cudaStream_t>(ov->side_stream));
#endif
  memset(ov, 0, sizeof(*ov));
  return DENO_OK;
}

// ------------------------------------------------------------------------------------------ deno_fno.h
int deno_fno_supported(const deno_fno_desc* desc) {
  FnoShape f;
  return parse_fno(desc, f) == DENO_OK ? 1 : 0;
}

size_t deno_fno_workspace_bytes(const deno_fno_desc* desc, int which) {
  FnoShape f;
  if (parse_fno(desc, f) != DENO_OK) return 0;
  if (which == DENO_FNO_WS_LIFT_BWD) {
    LiftBwdWs w;
    lift_bwd_ws(f, w);
    return w.total;
  }
  if (which == DENO_FNO_WS_PROJ_BWD) {
    ProjBwdWs w;
    if (proj_bwd_ws(f, w) != DENO_OK) return 0;
    return w.total;
  }
  fail(DENO_E_INVALID, "unknown workspace kind");
  return 0;
}

int deno_fno_lift_fwd(const deno_fno_desc* desc, const float* x, const float* grid, const float* w0, const float* b0,
                      float* h, void* stream) {
  g_launches = 0;
  FnoShape f;
  int rc = parse_fno(desc, f);
  if (rc != DENO_OK) return rc;
  if (!x || !w0 || !b0 || !h || (f.ND > 0 && !grid)) return fail(DENO_E_INVALID, "deno_fno_lift_fwd: null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LiftParams p;
  p.g = f.g; p.C = f.C; p.Kx = f.Kx; p.ND = f.ND;
  p.x = x; p.grid = grid; p.w0 = w0; p.b0 = b0; p.h = h;
  p.chunks = (f.g.nppts + kLiftThreads - 1) / kLiftThreads;
  const size_t smem = (size_t)f.C * (f.Kx + f.ND + 1) * sizeof(float);
  const int K = f.Kx + f.ND;
  prof_before("fno_lift", st);
#define DENO_LIFT_CASE(KT) DENO_LAUNCH(fno_lift_kernel<KT>, dim3((unsigned)(f.B * p.chunks)), dim3(kLiftThreads), smem, st, p)
  if (K <= 2) DENO_LIFT_CASE(2);
  else if (K <= 3) DENO_LIFT_CASE(3);
  else if (K <= 4) DENO_LIFT_CASE(4);
  else if (K <= 8) DENO_LIFT_CASE(8);
  else if (K <= 12) DENO_LIFT_CASE(12);
  else if (K <= 13) DENO_LIFT_CASE(13);
  else DENO_LIFT_CASE(16);
#undef DENO_LIFT_CASE
  return check_launch("fno_lift", st);
}

int deno_fno_lift_bwd(const deno_fno_desc* desc, const float* gh, const float* x, const float* grid, const float* w0,
                      float* gx, float* gw0, float* gb0, void* workspace, size_t workspace_bytes, void* stream) {
  g_launches = 0;
  FnoShape f;
  int rc = parse_fno(desc, f);
  if (rc != DENO_OK) return rc;
  if (!gh || !x || !w0 || (f.ND > 0 && !grid) || !gw0 || !gb0) return fail(DENO_E_INVALID, "deno_fno_lift_bwd: null pointer argument");
  LiftBwdWs w;
  lift_bwd_ws(f, w);
  if (workspace == nullptr || workspace_bytes < w.total) return fail(DENO_E_WORKSPACE, "deno_fno_lift_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LiftBwdParams p;
  p.g = f.g; p.C = f.C; p.Kx = f.Kx; p.ND = f.ND;
  p.gh = gh; p.x = x; p.grid = grid; p.w0 = w0; p.gx = gx;
  p.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.partial);
  p.chunks = (f.g.nppts + kLiftThreads - 1) / kLiftThreads;
  const size_t smem = lift_bwd_smem_bytes(f.C, f.Kx + f.ND);
  rc = allow_smem(fno_lift_bwd_kernel, smem, "fno_lift_bwd");
  if (rc != DENO_OK) return rc;
  prof_before("fno_lift_bwd", st);
  DENO_LAUNCH(fno_lift_bwd_kernel, dim3((unsigned)w.nblk), dim3(kLiftThreads), smem, st, p);
  rc = check_launch("fno_lift_bwd", st);
  if (rc != DENO_OK) return rc;
  return launch_reduce(p.partial, w.nblk, w.nj, w.nj, f.C * (f.Kx + f.ND), gw0, gb0, "fno_lift_reduce", st);
}

int deno_fno_proj_fwd(const deno_fno_desc* desc, const float* h, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* y, void* stream) {
  g_launches = 0;
  FnoShape f;
  int rc = parse_fno(desc, f);
  if (rc != DENO_OK) return rc;
  if (!h || !w1 || !b1 || !w2 || !b2 || !y) return fail(DENO_E_INVALID, "deno_fno_proj_fwd: null pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProjParams p;
  memset(&p, 0, sizeof(p));
  p.g = f.g; p.C = f.C; p.Hd = f.Hd; p.O = f.O;
  p.h = h; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.y = y;
#ifndef DENO_EMULATE
  {
    tc::ProjTcParams tp = proj_tc_params(f);
    tp.h = h; tp.w1 = w1; tp.b1 = b1; tp.w2 = w2; tp.b2 = b2; tp.y = y;
    rc = try_launch_proj_fwd_tc(f, tp, st);
    if (rc != -1) return rc;
  }
#endif
  const int per = kProjThreads * proj_fwd_pt(f.C);
  p.chunks = (f.g.nppts + per - 1) / per;
  const size_t smem = proj_fwd_smem_bytes(f.C, f.Hd, f.O);
  if (smem > (size_t)kMaxSmemBytes) return fail(DENO_E_UNSUPPORTED, "projection weights do not fit shared memory");
#define DENO_PROJ_FWD(CV)                                                                                      \
  {                                                                                                            \
    auto kern = fno_proj_fwd_kernel<CV, proj_fwd_pt(CV)>;                                                      \
    rc = allow_smem(kern, smem, "fno_proj_fwd");                                                               \
    if (rc != DENO_OK) return rc;                                                                              \
    prof_before("fno_proj_fwd", st);                                                                           \
    DENO_LAUNCH(kern, dim3((unsigned)(f.B * p.chunks)), dim3(kProjThreads), smem, st, p);                      \
  }
  DENO_PROJ_DISPATCH(f.C, DENO_PROJ_FWD)
#undef DENO_PROJ_FWD
  return check_launch("fno_proj_fwd", st);
}

int deno_fno_proj_bwd(const deno_fno_desc* desc, const float* gy, const float* h, const float* w1, const float* b1,
                      const float* w2, float* gh, float* gw1, float* gb1, float* gw2, float* gb2, void* workspace,
                      size_t workspace_bytes, void* stream) {
  g_launches = 0;
  FnoShape f;
  int rc = parse_fno(desc, f);
  if (rc != DENO_OK) return rc;
  if (!gy || !h || !w1 || !b1 || !w2 || !gh || !gw1 || !gb1 || !gw2 || !gb2)
    return fail(DENO_E_INVALID, "deno_fno_proj_bwd: null pointer argument");
  ProjBwdWs w;
  rc = proj_bwd_ws(f, w);
  if (rc != DENO_OK) return rc;
  if (workspace == nullptr || workspace_bytes < w.total) return fail(DENO_E_WORKSPACE, "deno_fno_proj_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  ProjParams p;
  memset(&p, 0, sizeof(p));
  p.g = f.g; p.C = f.C; p.Hd = f.Hd; p.O = f.O;
  p.h = h; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = nullptr; p.gy = gy; p.gh = gh;
  p.gt = reinterpret_cast<float*>(ws + w.gt);
  p.partial = reinterpret_cast<float*>(ws + w.small);
  int nparts = w.nblk;
  bool done = false;
#ifndef DENO_EMULATE
  if (proj_bwd_tc_eligible(f)) {
    tc::ProjTcParams tp = proj_tc_params(f);
    tp.h = h; tp.w1 = w1; tp.b1 = b1; tp.w2 = w2; tp.gy = gy; tp.gh = gh; tp.gt = p.gt; tp.partial = p.partial;
    rc = launch_proj_bwd_tc(tp, nparts, st);
    if (rc != DENO_OK) return rc;
    done = true;
  }
#endif
  const int per = kProjThreads * proj_bwd_pt(f.C);
  p.chunks = (f.g.nppts + per - 1) / per;
  const size_t smem = proj_bwd_smem_bytes(f.C, f.Hd, f.O);
  if (!done && smem > (size_t)kMaxSmemBytes) return fail(DENO_E_UNSUPPORTED, "projection weights do not fit shared memory");
#define DENO_PROJ_BWD(CV)                                                                                      \
  {                                                                                                            \
    auto kern = fno_proj_bwd_kernel<CV, proj_bwd_pt(CV)>;                                                      \
    rc = allow_smem(kern, smem, "fno_proj_bwd");                                                               \
    if (rc != DENO_OK) return rc;                                                                              \
    prof_before("fno_proj_bwd", st);                                                                           \
    DENO_LAUNCH(kern, dim3((unsigned)w.nblk), dim3(kProjThreads), smem, st, p);                                \
  }
  if (!done) {
    DENO_PROJ_DISPATCH(f.C, DENO_PROJ_BWD)
    rc = check_launch("fno_proj_bwd", st);
    if (rc != DENO_OK) return rc;
  }
#undef DENO_PROJ_BWD
  // a partial row is gb1 | gw2 | gb2: two passes of the reduce kernel ("first nw columns / the rest") split it
  rc = launch_reduce(p.partial, nparts, w.nsmall, f.Hd + f.O * f.Hd, f.Hd, gb1, gw2, "fno_proj_reduce", st);
  if (rc != DENO_OK) return rc;
  rc = launch_reduce(p.partial + f.Hd + f.O * f.Hd, nparts, w.nsmall, f.O, f.O, gb2, nullptr, "fno_proj_reduce", st);
  if (rc != DENO_OK) return rc;
  const Shape s = proj_wgrad_shape(f);
  float* wpart = reinterpret_cast<float*>(ws + w.wpart);
#ifndef DENO_EMULATE
  if (done) {                                          // gt is blocked: tensor-core weight gradient only
    int np = try_launch_bypass_wgrad_tma(s, p.gt, h, wpart, st, (f.g.nppts + 127) / 128);
    if (np == -1) np = try_launch_bypass_wgrad_tc(s, p.gt, h, wpart, st, (f.g.nppts + 127) / 128);
    if (np <= 0) return np == -2 ? DENO_E_CUDA : fail(DENO_E_UNSUPPORTED, "bypass_wgrad_tc refused the projection shape");
    return launch_reduce(wpart, np, f.Hd * f.C, f.Hd * f.C, f.Hd * f.C, gw1, nullptr, "bypass_wgrad_reduce", st);
  }
#endif
  return launch_bypass_wgrad(s, p.gt, h, nullptr, wpart, gw1, nullptr, true, st);
}

}  // extern "C"
