// spectral_simt.cuh - fp32 SIMT kernels of the spectral-convolution path (all shapes).
//
// These are the general kernels: any grid size, any mode count, any channel count.  They
// implement the truncated-DFT statement of the layer (oracle/dft_truth.py) stage by stage:
//
//   plane_analysis   x (or gy*act'(z))  -> packed plane spectrum        (forward truncated rDFT,
//                                          replaces the full rfftn of spectral_layers.py:95,191,310)
//   outer_analysis   3-D only: contracts the leading axis X
//   mode_mix         per-mode complex channel product                   (compl_mul*, :73-82,172-181,287-297)
//   mode_mix_wgrad   weight gradient of the product
//   outer_synthesis  3-D only: expands the leading axis X
//   plane_synthesis  zero-padded inverse transform fused with the 1x1 bypass conv, bias and the
//                    activation (replaces zeros + scatter + irfftn + add + act, :96-110,194-203,311-333)
//   bypass_wgrad     gradient of the 1x1 conv weight / bias (two-stage deterministic reduction)
//
// Twiddle tables are produced on the host in fp64 (capi.cu) and laid out padded:
//   twW[y][MWp]  (cos, sin)(2 pi l y / W),   MWp = round_up(MW, 4), zero in the padding
//   twH[x][KH4]  (cos, sin)(2 pi f_k x / H), KH4 = round_up(KH, 4), zero in the padding
//   cl[MWp]      Hermitian weights of the C2R transform
#pragma once
#include "spectral_common.cuh"

namespace deno {

constexpr int kSynthMaxRows = 8;  // rows per CTA tile in plane_synthesis (register accumulators)

// ----------------------------------------------------------------------------------------------
// plane_analysis
// ----------------------------------------------------------------------------------------------
struct AnalysisParams {
  PlaneGeom g;
  const float* u;      // x, or gy when FUSE_GZ
  const float* z;      // FUSE_GZ: saved activation derivative act'(z)
  float* gz_out;       // FUSE_GZ: gz = gy * act'(z) is written here (same geometry as u)
  float2* out;         // [nplanes][KH][MW]
  const float2* twW;
  const float2* twH;
  const float* cl;
  float scale;
  int herm;            // multiply mode l by cl[l]
  int act;
  int nplanes;         // NBO*NBI*C
  int G;               // planes per CTA
  int RT;              // rows per tile (<= RTP)
  int RTP;             // thread mapping: r = tid % RTP, lq = tid / RTP
  int YC;              // columns per chunk
  int RS;              // row split of stage 2
  int CS;              // column split of stage 1: thread = (row, mode quad, column slice); 1 except for few-row tiles
};

struct AnalysisSmem {
  int xs, tws, ts, twh, acc, total;  // offsets / total in floats
};

__host__ __device__ inline AnalysisSmem analysis_smem(const AnalysisParams& p) {
  const int MWp = round_up(p.g.MW, 4), KH4 = round_up(p.g.KH, 4);
  AnalysisSmem s;
  int off = 0;
  s.xs = off;  off += round_up(p.RT * (p.YC + 1), 4);
  s.tws = off; off += p.YC * MWp * 2;
  s.ts = off;  off += p.RT * MWp * 2 * (p.CS > 1 ? p.CS : 1);   // per column slice, reduced in place
  s.twh = off; off += p.g.H * KH4 * 2;
  s.acc = off; off += p.RS * p.G * KH4 * p.g.MW * 2;
  s.total = off;
  return s;
}

template <bool FUSE_GZ>
__global__ void plane_analysis_kernel(AnalysisParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int H = p.g.H, W = p.g.W, KH = p.g.KH, MW = p.g.MW;
  const int MWp = round_up(MW, 4), KH4 = round_up(KH, 4), NLG = MWp / 4, KQ = KH4 / 4;
  const int YCp = p.YC + 1;
  const AnalysisSmem so = analysis_smem(p);
  float* xs = smem + so.xs;
  float2* tws = reinterpret_cast<float2*>(smem + so.tws);
  float2* Ts = reinterpret_cast<float2*>(smem + so.ts);
  float2* twHs = reinterpret_cast<float2*>(smem + so.twh);
  float2* acc2 = reinterpret_cast<float2*>(smem + so.acc);

  const int q0 = blockIdx.x * p.G;
  const int Gc = min(p.G, p.nplanes - q0);
  const int rows = Gc * H;
  const int acc_stride = p.G * KH4 * MW;  // per row-split slice

  for (int i = tid; i < H * KH4; i += nt) twHs[i] = p.twH[i];
  for (int i = tid; i < p.RS * acc_stride; i += nt) acc2[i] = make_float2(0.f, 0.f);

  const int r = tid % p.RTP, lq = (tid / p.RTP) % NLG, cs = tid / (p.RTP * NLG);   // cs < CS (launch geometry)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int ntask2 = p.G * KQ * MW;

  for (int tile0 = 0; tile0 < rows; tile0 += p.RT) {
    const int trows = min(p.RT, rows - tile0);
    const bool active = (r < trows) && (cs < p.CS);
    float re0 = 0.f, re1 = 0.f, re2 = 0.f, re3 = 0.f, im0 = 0.f, im1 = 0.f, im2 = 0.f, im3 = 0.f;

    for (int y0 = 0; y0 < W; y0 += p.YC) {
      const int yc = min(p.YC, W - y0);
      __syncthreads();  // previous chunk / previous tile fully consumed (also orders the init above)
      for (int rr = warp; rr < trows; rr += nwarps) {
        const int row = tile0 + rr;
        const int gi = row / H, xr = row - gi * H;
        const int q = q0 + gi;
        const long long rbase = plane_base(p.g, q / p.g.C, q % p.g.C) + (long long)xr * W + y0;
        for (int j = lane; j < yc; j += 32) {
          float v = p.u[rbase + j];
          if (FUSE_GZ) {
            v *= p.z[rbase + j];
            p.gz_out[rbase + j] = v;
          }
          xs[rr * YCp + j] = v;
        }
      }
      for (int idx = tid; idx < yc * MWp; idx += nt) tws[idx] = p.twW[(long long)y0 * MWp + idx];
      __syncthreads();
      if (active) {
        const float* xrow = xs + r * YCp;
        const float4* t4 = reinterpret_cast<const float4*>(tws) + lq * 2;
        const int step = MWp / 2;
#pragma unroll 4
        for (int j = cs; j < yc; j += p.CS) {
          const float xv = xrow[j];
          const float4 a = t4[j * step], b = t4[j * step + 1];
          re0 = fmaf(xv, a.x, re0); im0 = fmaf(-xv, a.y, im0);
          re1 = fmaf(xv, a.z, re1); im1 = fmaf(-xv, a.w, im1);
          re2 = fmaf(xv, b.x, re2); im2 = fmaf(-xv, b.y, im2);
          re3 = fmaf(xv, b.z, re3); im3 = fmaf(-xv, b.w, im3);
        }
      }
    }
    if (active) {
      float2* t = Ts + (cs * p.RT + r) * MWp + lq * 4;
      t[0] = make_float2(re0, im0);
      t[1] = make_float2(re1, im1);
      t[2] = make_float2(re2, im2);
      t[3] = make_float2(re3, im3);
    }
    __syncthreads();
    if (p.CS > 1) {                       // sum the column slices into slice 0 (fixed order)
      for (int i = tid; i < trows * MWp; i += nt) {
        float2 a = Ts[i];
        for (int c2 = 1; c2 < p.CS; ++c2) {
          const float2 b = Ts[c2 * p.RT * MWp + i];
          a.x += b.x;
          a.y += b.y;
        }
        Ts[i] = a;
      }
      __syncthreads();
    }
    // stage 2: contract the rows of each plane against e^{-i phi(k, x)}; thread = (row split, plane, k quad, l)
    for (int tt = tid; tt < ntask2 * p.RS; tt += nt) {
      const int rs = tt / ntask2, task = tt - rs * ntask2;
      const int l = task % MW;
      const int kq = (task / MW) % KQ;
      const int gi = task / (MW * KQ);
      if (gi >= Gc) continue;
      const int lo = max(gi * H, tile0), hi = min((gi + 1) * H, tile0 + trows);
      float ar[4] = {0.f, 0.f, 0.f, 0.f}, ai[4] = {0.f, 0.f, 0.f, 0.f};
      for (int row = lo + rs; row < hi; row += p.RS) {
        const float2 t = Ts[(row - tile0) * MWp + l];
        const float4* e4 = reinterpret_cast<const float4*>(twHs + (row - gi * H) * KH4 + kq * 4);
        const float4 ea = e4[0], eb = e4[1];
        // (c - i s) * (tr + i ti) = (c tr + s ti) + i (c ti - s tr)
        ar[0] = fmaf(ea.x, t.x, fmaf(ea.y, t.y, ar[0])); ai[0] = fmaf(ea.x, t.y, fmaf(-ea.y, t.x, ai[0]));
        ar[1] = fmaf(ea.z, t.x, fmaf(ea.w, t.y, ar[1])); ai[1] = fmaf(ea.z, t.y, fmaf(-ea.w, t.x, ai[1]));
        ar[2] = fmaf(eb.x, t.x, fmaf(eb.y, t.y, ar[2])); ai[2] = fmaf(eb.x, t.y, fmaf(-eb.y, t.x, ai[2]));
        ar[3] = fmaf(eb.z, t.x, fmaf(eb.w, t.y, ar[3])); ai[3] = fmaf(eb.z, t.y, fmaf(-eb.w, t.x, ai[3]));
      }
      float2* a = acc2 + rs * acc_stride + (gi * KH4 + kq * 4) * MW + l;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 v = a[j * MW];
        v.x += ar[j];
        v.y += ai[j];
        a[j * MW] = v;
      }
    }
  }
  __syncthreads();
  for (int o = tid; o < Gc * KH * MW; o += nt) {
    const int gi = o / (KH * MW);
    const int rem = o - gi * KH * MW;
    const int k = rem / MW, l = rem - k * MW;
    float sr = 0.f, si = 0.f;
    for (int rs = 0; rs < p.RS; ++rs) {
      const float2 v = acc2[rs * acc_stride + (gi * KH4 + k) * MW + l];
      sr += v.x;
      si += v.y;
    }
    const float f = p.scale * (p.herm ? p.cl[l] : 1.0f);
    p.out[(long long)(q0 + gi) * KH * MW + rem] = make_float2(sr * f, si * f);
  }
}

// ----------------------------------------------------------------------------------------------
// line_analysis: the 1-D layers' truncated forward rDFT (and, FUSE_GZ, gz = gy * act'(z) on the way), one (b, c) line
// of W points per "plane".  plane_analysis_kernel is built around planes with many rows; for single-row planes it left
// a CTA with 8 rows x 4 mode quads of parallelism walking ~1 000 columns (41 / 45 us at the Burgers shape, 5 MB of
// input).  Here a CTA stages kLineRows lines in shared memory and EVERY thread owns one retained mode l and one of NS
// interleaved column slices for all staged lines: the twiddle (one coalesced global load per column, L2-resident
// table) is reused across the lines, the line values are shared-memory broadcasts, and the slices are summed in a
// fixed order at the end (bitwise reproducible).
// ----------------------------------------------------------------------------------------------
constexpr int kLineRows = 8;
constexpr int kLineThreads = 256;

struct LineAnalysisParams {
  PlaneGeom g;           // H == 1
  const float* u;        // x, or gy when FUSE_GZ
  const float* z;
  float* gz_out;
  float2* out;           // [nlines][MW]
  const float2* twW;     // [W][MWp] (cos, sin)
  const float* cl;
  float scale;
  int herm;
  int nlines;            // NBO * NBI * C
  int MWp;               // retained modes rounded up to 4 (twiddle pitch); MWp <= kLineThreads
  int NS;                // column slices = kLineThreads / MWp
};

__host__ __device__ inline size_t line_analysis_smem(int W, int MWp, int NS) {
  return ((size_t)kLineRows * (W + 1) + (size_t)NS * kLineRows * MWp * 2) * sizeof(float);
}

template <bool FUSE_GZ>
__global__ void __launch_bounds__(kLineThreads) line_analysis_kernel(LineAnalysisParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* xs = reinterpret_cast<float*>(smem_raw);                         // [kLineRows][W + 1]
  const int W = p.g.W, MW = p.g.MW, MWp = p.MWp, NS = p.NS, Wp = W + 1;
  float2* red = reinterpret_cast<float2*>(xs + kLineRows * Wp);           // [NS][kLineRows][MWp]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kLineThreads / 32;
  const int q0 = blockIdx.x * kLineRows;
  const int nl = min(kLineRows, p.nlines - q0);

  for (int rr = warp; rr < kLineRows; rr += nwarps) {                      // one warp per line: coalesced
    float* dst = xs + rr * Wp;
    if (rr < nl) {
      const int q = q0 + rr;
      const long long base = plane_base(p.g, q / p.g.C, q % p.g.C);
      // eight independent loads (sixteen with act'(z)) in flight per lane: a loop of dependent load -> store pairs cost
      // one memory round trip per 32 columns (35 us for the gz variant at the Burgers shape)
      for (int j0 = 0; j0 < W; j0 += 8 * 32) {
        float v[8], zz[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int j = min(j0 + t * 32 + lane, W - 1);
          v[t] = ldg_ordered(p.u + base + j);
          if (FUSE_GZ) zz[t] = ldg_ordered(p.z + base + j);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int j = j0 + t * 32 + lane;
          if (j < W) {
            float val = v[t];
            if (FUSE_GZ) {
              val *= zz[t];
              p.gz_out[base + j] = val;
            }
            dst[j] = val;
          }
        }
      }
    } else {
      for (int j = lane; j < W; j += 32) dst[j] = 0.f;
    }
  }
  __syncthreads();

  const int l = tid % MWp, sl = tid / MWp;
  if (sl < NS) {
    float re[kLineRows], im[kLineRows];
#pragma unroll
    for (int r = 0; r < kLineRows; ++r) re[r] = im[r] = 0.f;
    const float2* tw = p.twW + l;
    for (int y0 = sl; y0 < W; y0 += 4 * NS) {                              // four twiddle loads in flight
      float2 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int y = y0 + u * NS;
        t[u] = y < W ? tw[(long long)y * MWp] : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int y = min(y0 + u * NS, W - 1);                             // t[u] = 0 beyond the line
#pragma unroll
        for (int r = 0; r < kLineRows; ++r) {
          const float xv = xs[r * Wp + y];
          re[r] = fmaf(xv, t[u].x, re[r]);
          im[r] = fmaf(-xv, t[u].y, im[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kLineRows; ++r) red[(sl * kLineRows + r) * MWp + l] = make_float2(re[r], im[r]);
  }
  __syncthreads();
  for (int o = tid; o < nl * MW; o += kLineThreads) {
    const int r = o / MW, lo = o - r * MW;
    float sr = 0.f, si = 0.f;
    for (int s2 = 0; s2 < NS; ++s2) {
      const float2 v = red[(s2 * kLineRows + r) * MWp + lo];
      sr += v.x;
      si += v.y;
    }
    const float f = p.scale * (p.herm ? p.cl[lo] : 1.0f);
    p.out[(long long)(q0 + r) * MW + lo] = make_float2(sr * f, si * f);
  }
}

// ----------------------------------------------------------------------------------------------
// outer axis (3-D): in/out are complex arrays indexed [b][x or k][c][Q] / [b][c][k][Q]
// ----------------------------------------------------------------------------------------------
struct OuterParams {
  const float2* in;
  float2* out;
  const float2* twX;  // [X][KX4]
  int B, C, X, KX, Q;
};

// out[b][c][k][q] = sum_x (cos - i sin)(k,x) * in[b][x][c][q];  thread = (q, k quad, c, b)
__global__ void outer_analysis_kernel(OuterParams p) {
  DENO_PDL_SYNC();
  const int KX4 = round_up(p.KX, 4), KQ = KX4 / 4;
  const long long total = (long long)p.B * p.C * KQ * p.Q;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int q = int(t % p.Q);
    long long rest = t / p.Q;
    const int kq = int(rest % KQ);
    rest /= KQ;
    const int c = int(rest % p.C);
    const int b = int(rest / p.C);
    float ar[4] = {0.f, 0.f, 0.f, 0.f}, ai[4] = {0.f, 0.f, 0.f, 0.f};
    const float2* src = p.in + ((long long)b * p.X * p.C + c) * p.Q + q;
    for (int x = 0; x < p.X; ++x) {
      const float2 v = src[(long long)x * p.C * p.Q];
      const float4* e4 = reinterpret_cast<const float4*>(p.twX + x * KX4 + kq * 4);
      const float4 ea = e4[0], eb = e4[1];
      ar[0] = fmaf(ea.x, v.x, fmaf(ea.y, v.y, ar[0])); ai[0] = fmaf(ea.x, v.y, fmaf(-ea.y, v.x, ai[0]));
      ar[1] = fmaf(ea.z, v.x, fmaf(ea.w, v.y, ar[1])); ai[1] = fmaf(ea.z, v.y, fmaf(-ea.w, v.x, ai[1]));
      ar[2] = fmaf(eb.x, v.x, fmaf(eb.y, v.y, ar[2])); ai[2] = fmaf(eb.x, v.y, fmaf(-eb.y, v.x, ai[2]));
      ar[3] = fmaf(eb.z, v.x, fmaf(eb.w, v.y, ar[3])); ai[3] = fmaf(eb.z, v.y, fmaf(-eb.w, v.x, ai[3]));
    }
    float2* dst = p.out + (((long long)b * p.C + c) * p.KX + kq * 4) * p.Q + q;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (kq * 4 + j < p.KX) dst[(long long)j * p.Q] = make_float2(ar[j], ai[j]);
  }
}

// out[b][x][c][q] = sum_k (cos + i sin)(k,x) * in[b][c][k][q];  thread = (q, x quad, c, b)
__global__ void outer_synthesis_kernel(OuterParams p) {
  DENO_PDL_SYNC();
  const int KX4 = round_up(p.KX, 4);
  const int XQ = (p.X + 3) / 4;
  const long long total = (long long)p.B * p.C * XQ * p.Q;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int q = int(t % p.Q);
    long long rest = t / p.Q;
    const int xq = int(rest % XQ);
    rest /= XQ;
    const int c = int(rest % p.C);
    const int b = int(rest / p.C);
    float ar[4] = {0.f, 0.f, 0.f, 0.f}, ai[4] = {0.f, 0.f, 0.f, 0.f};
    const float2* src = p.in + (((long long)b * p.C + c) * p.KX) * p.Q + q;
    for (int k = 0; k < p.KX; ++k) {
      const float2 v = src[(long long)k * p.Q];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = min(xq * 4 + j, p.X - 1);
        const float2 e = p.twX[x * KX4 + k];
        // (c + i s) * (vr + i vi)
        ar[j] = fmaf(e.x, v.x, fmaf(-e.y, v.y, ar[j]));
        ai[j] = fmaf(e.x, v.y, fmaf(e.y, v.x, ai[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = xq * 4 + j;
      if (x < p.X) p.out[(((long long)b * p.X + x) * p.C + c) * p.Q + q] = make_float2(ar[j], ai[j]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// mode_mix:  out[b,co,p] = sum_ci in[b,ci,p] * (conj?) W[ci,co,p]
// ----------------------------------------------------------------------------------------------
struct MixParams {
  const float2* in;
  float2* out;
  const float2* w[4];
  ModeDecode md;
  int B, CI, CO;
  long long w_sci, w_sco;  // complex-element strides of the contracted / produced channel inside a block
  int conj_w;
};

constexpr int kMixBatchTile = 1;   // one batch entry per thread: the kernel is latency-bound, so favour more warps
constexpr int kMixUnroll = 8;

// blockDim = (32 modes, 8 output channels); grid = (ceil(M/32), ceil(CO/8), ceil(B/4))
__device__ __forceinline__ void mode_mix_body(const MixParams& p, int bx, int by, int bz) {
  const int M = p.md.M;
  const int pm = bx * blockDim.x + threadIdx.x;
  const int co = by * blockDim.y + threadIdx.y;
  const int b0 = bz * kMixBatchTile;
  if (pm >= M || co >= p.CO) return;
  int blk, off;
  decode_mode(p.md, pm, blk, off);
  const float2* wb = blk == 0 ? p.w[0] : (blk == 1 ? p.w[1] : (blk == 2 ? p.w[2] : p.w[3]));
  const float2* w = wb + (long long)co * p.w_sco + off;
  float ar[kMixBatchTile], ai[kMixBatchTile];
#pragma unroll
  for (int j = 0; j < kMixBatchTile; ++j) ar[j] = ai[j] = 0.f;
  const float sgn = p.conj_w ? -1.0f : 1.0f;
  const float2* inb[kMixBatchTile];
#pragma unroll
  for (int j = 0; j < kMixBatchTile; ++j) inb[j] = p.in + ((long long)min(b0 + j, p.B - 1) * p.CI) * M + pm;
  // The loop is latency-bound (every operand comes from L2), so all loads of a batch of kMixUnroll
  // contracted channels are issued before any of them is consumed.
  for (int ci0 = 0; ci0 < p.CI; ci0 += kMixUnroll) {
    float2 wv[kMixUnroll], v[kMixUnroll][kMixBatchTile];
#pragma unroll
    for (int u = 0; u < kMixUnroll; ++u) {
      const int ci = min(ci0 + u, p.CI - 1);
      wv[u] = ldg_ordered(w + (long long)ci * p.w_sci);
#pragma unroll
      for (int j = 0; j < kMixBatchTile; ++j) v[u][j] = ldg_ordered(inb[j] + (long long)ci * M);
    }
    DENO_SCHED_FENCE();   // keep the whole batch of loads ahead of the arithmetic (the compiler otherwise re-serialises them)
#pragma unroll
    for (int u = 0; u < kMixUnroll; ++u) {
      const float wr = (ci0 + u < p.CI) ? wv[u].x : 0.f;
      const float wi = (ci0 + u < p.CI) ? wv[u].y * sgn : 0.f;
#pragma unroll
      for (int j = 0; j < kMixBatchTile; ++j) {
        ar[j] = fmaf(v[u][j].x, wr, fmaf(-v[u][j].y, wi, ar[j]));
        ai[j] = fmaf(v[u][j].x, wi, fmaf(v[u][j].y, wr, ai[j]));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMixBatchTile; ++j)
    if (b0 + j < p.B) p.out[((long long)(b0 + j) * p.CO + co) * M + pm] = make_float2(ar[j], ai[j]);
}

__global__ void mode_mix_kernel(MixParams p) {
  DENO_PDL_SYNC();
  mode_mix_body(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// ----------------------------------------------------------------------------------------------
// mode_mix_lane: the same product as mode_mix (out[b,co,p] = sum_ci in[b,ci,p] * (conj?) W[ci,co,p]) for layers
// where the contracted width makes the weight stream the cost (width >= 64: 9.4 - 37.7 MB of weights per layer).
// Lane = packed mode, exactly like mode_mix (every global access is a coalesced 256-byte row of the mode-innermost
// layouts), but a thread owns kMixLaneB batch entries of one produced channel: the weight element it loads is used
// kMixLaneB times, and the contracted-channel spectrum of those batch entries is staged in shared memory once per
// block instead of being re-read from L2 by every produced channel.  Weights cross L2 -> SM  B / kMixLaneB  times
// instead of B times; no operand images, no tensor cores: the arithmetic is 0.2 GFLOP.
// blockDim = (32 modes, 8 produced channels); grid = (ceil(M / 32), ceil(CO / 8), ceil(B / kMixLaneB)).
// ----------------------------------------------------------------------------------------------
constexpr int kMixLaneB = 5;        // batch entries per thread
constexpr int kMixLaneChunk = 64;   // contracted channels staged per pass
constexpr int kMixLaneUnroll = 8;   // weight loads in flight per thread (16 measured slower: C3 layer 280 -> 288 us, C4b 1822 -> 1848)

__host__ __device__ inline size_t mix_lane_smem_bytes(int CI) {
  const int cc = CI < kMixLaneChunk ? CI : kMixLaneChunk;
  return (size_t)kMixLaneB * cc * 32 * sizeof(float2);
}

__global__ void __launch_bounds__(256) mode_mix_lane_kernel(MixParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(ml_smem);
  float2* xs = reinterpret_cast<float2*>(ml_smem);              // [kMixLaneB][chunk][32]
  const int M = p.md.M;
  const int lane = threadIdx.x, cy = threadIdx.y;
  const int tid = cy * 32 + lane;
  const int pm = blockIdx.x * 32 + lane;
  const int pmc = min(pm, M - 1);
  const int co = blockIdx.y * 8 + cy;
  const int b0 = blockIdx.z * kMixLaneB;
  const int nb = min(kMixLaneB, p.B - b0);
  const int cc = p.CI < kMixLaneChunk ? p.CI : kMixLaneChunk;
  int blk, off;
  decode_mode(p.md, pmc, blk, off);
  const float2* wb = blk == 0 ? p.w[0] : (blk == 1 ? p.w[1] : (blk == 2 ? p.w[2] : p.w[3]));
  const float2* w = wb + (long long)min(co, p.CO - 1) * p.w_sco + off;
  const float sgn = p.conj_w ? -1.0f : 1.0f;
  float ar[kMixLaneB], ai[kMixLaneB];
#pragma unroll
  for (int j = 0; j < kMixLaneB; ++j) ar[j] = ai[j] = 0.f;

  for (int c0 = 0; c0 < p.CI; c0 += cc) {
    const int ncc = min(cc, p.CI - c0);
    if (c0 > 0) __syncthreads();                                  // previous chunk consumed
    // stage in[b0 .. b0 + nb)[c0 .. c0 + ncc)[32 modes]: rows of 256 bytes, eight loads in flight per thread
    const int nrow = nb * ncc;
    for (int r0 = cy; r0 < nrow; r0 += 8 * 8) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = min(r0 + u * 8, nrow - 1);
        const int j = r / ncc, ci = r - j * ncc;
        v[u] = ldg_ordered(p.in + ((long long)(b0 + j) * p.CI + c0 + ci) * M + pmc);
      }
      DENO_SCHED_FENCE();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = r0 + u * 8;
        if (r < nrow) {
          const int j = r / ncc, ci = r - j * ncc;
          xs[(j * cc + ci) * 32 + lane] = v[u];
        }
      }
    }
    __syncthreads();
    const float2* wc = w + (long long)c0 * p.w_sci;
    for (int ci0 = 0; ci0 < ncc; ci0 += kMixLaneUnroll) {
      float2 wv[kMixLaneUnroll];
#pragma unroll
      for (int u = 0; u < kMixLaneUnroll; ++u) wv[u] = ldg_ordered(wc + (long long)min(ci0 + u, ncc - 1) * p.w_sci);
      DENO_SCHED_FENCE();
#pragma unroll
      for (int u = 0; u < kMixLaneUnroll; ++u) {
        const bool live = ci0 + u < ncc;
        const float wr = live ? wv[u].x : 0.f;
        const float wi = live ? wv[u].y * sgn : 0.f;
        const float2* xp = xs + min(ci0 + u, ncc - 1) * 32 + lane;
#pragma unroll
        for (int j = 0; j < kMixLaneB; ++j) {
          const float2 x = xp[j * cc * 32];                      // rows of batch entries >= nb hold stale data: never stored
          ar[j] = fmaf(x.x, wr, fmaf(-x.y, wi, ar[j]));
          ai[j] = fmaf(x.x, wi, fmaf(x.y, wr, ai[j]));
        }
      }
    }
  }
  (void)tid;
  if (pm < M && co < p.CO) {
#pragma unroll
    for (int j = 0; j < kMixLaneB; ++j)
      if (j < nb) p.out[((long long)(b0 + j) * p.CO + co) * M + pm] = make_float2(ar[j], ai[j]);
  }
}

// gW[i,o,p] = sum_b conj(xhat[b,i,p]) * ghat[b,o,p]
struct MixWgradParams {
  const float2* xhat;
  const float2* ghat;
  float2* gw[4];
  ModeDecode md;
  int B, CI, CO;
};

constexpr int kMixWgradOut = 1;   // output channels per thread (latency-bound: favour more warps)

// blockDim = (32 modes, 8 output channels); grid = (ceil(M/32), ceil(CO/8), CI)
__device__ __forceinline__ void mode_mix_wgrad_body(const MixWgradParams& p, int bx, int by, int bz) {
  const int M = p.md.M;
  const int pm = bx * blockDim.x + threadIdx.x;
  const int oq = by * blockDim.y + threadIdx.y;
  const int ci = bz;
  if (pm >= M || oq * kMixWgradOut >= p.CO) return;
  int blk, off;
  decode_mode(p.md, pm, blk, off);
  float ar[kMixWgradOut], ai[kMixWgradOut];
#pragma unroll
  for (int j = 0; j < kMixWgradOut; ++j) ar[j] = ai[j] = 0.f;
  const float2* xp = p.xhat + (long long)ci * M + pm;
  const float2* gp[kMixWgradOut];
#pragma unroll
  for (int j = 0; j < kMixWgradOut; ++j) gp[j] = p.ghat + (long long)min(oq * kMixWgradOut + j, p.CO - 1) * M + pm;
  for (int b0 = 0; b0 < p.B; b0 += 4) {
    float2 xv[4], gv[4][kMixWgradOut];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int b = min(b0 + u, p.B - 1);
      xv[u] = ldg_ordered(xp + (long long)b * p.CI * M);
#pragma unroll
      for (int j = 0; j < kMixWgradOut; ++j) gv[u][j] = ldg_ordered(gp[j] + (long long)b * p.CO * M);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xr = (b0 + u < p.B) ? xv[u].x : 0.f, xi = (b0 + u < p.B) ? xv[u].y : 0.f;
#pragma unroll
      for (int j = 0; j < kMixWgradOut; ++j) {
        // conj(x) * g = (xr gr + xi gi) + i (xr gi - xi gr)
        ar[j] = fmaf(xr, gv[u][j].x, fmaf(xi, gv[u][j].y, ar[j]));
        ai[j] = fmaf(xr, gv[u][j].y, fmaf(-xi, gv[u][j].x, ai[j]));
      }
    }
  }
  float2* gb = blk == 0 ? p.gw[0] : (blk == 1 ? p.gw[1] : (blk == 2 ? p.gw[2] : p.gw[3]));
  float2* dst = gb + ((long long)ci * p.CO) * p.md.Mblk + off;
#pragma unroll
  for (int j = 0; j < kMixWgradOut; ++j)
    if (oq * kMixWgradOut + j < p.CO) dst[(long long)(oq * kMixWgradOut + j) * p.md.Mblk] = make_float2(ar[j], ai[j]);
}

__global__ void mode_mix_wgrad_kernel(MixWgradParams p) {
  DENO_PDL_SYNC();
  mode_mix_wgrad_body(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// One launch for the three small consumers of the packed gradient spectrum G^ in the backward pass:
//   blocks [0, nA)        G^x = conj(W) G^            (mode_mix_body)
//   blocks [nA, nA + nB)  gW  = sum_b conj(X^) G^     (mode_mix_wgrad_body)
//   block  nA + nB        bias gradient from the DC mode: glb[o] = S * sum_b Re G^[b, o, 0]
// Each of them is latency-bound and far too small to fill the GPU on its own; as separate launches they
// cost ~8 us each at the Darcy shape.
struct MixBwdFusedParams {
  MixParams mix;
  MixWgradParams wg;
  int ax, ay, az;      // grid of the mix role (0 blocks when gx is not needed)
  int bx, by, bz;      // grid of the weight-gradient role
  const float2* ghat;  // bias role (glb == nullptr: disabled)
  float* glb;
  int B, CO, M;
  float s_total;
};

__global__ void mode_mix_bwd_fused_kernel(MixBwdFusedParams f) {
  DENO_PDL_SYNC();
  int id = blockIdx.x;
  const int nA = f.ax * f.ay * f.az, nB = f.bx * f.by * f.bz;
  if (id < nA) {
    const int x = id % f.ax, y = (id / f.ax) % f.ay, z = id / (f.ax * f.ay);
    mode_mix_body(f.mix, x, y, z);
    return;
  }
  id -= nA;
  if (id < nB) {
    const int x = id % f.bx, y = (id / f.bx) % f.by, z = id / (f.bx * f.by);
    mode_mix_wgrad_body(f.wg, x, y, z);
    return;
  }
  if (f.glb != nullptr) {
    for (int o = threadIdx.y * blockDim.x + threadIdx.x; o < f.CO; o += blockDim.x * blockDim.y) {
      float acc = 0.f;
      for (int b0 = 0; b0 < f.B; b0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = ldg_ordered(reinterpret_cast<const float*>(f.ghat + ((long long)min(b0 + u, f.B - 1) * f.CO + o) * f.M));
          if (b0 + u >= f.B) v[u] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      f.glb[o] = acc * f.s_total;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Tiled per-mode products.  mode_mix / mode_mix_bwd_fused above give every (mode, channel, batch) output its own
// thread and read both operands straight from L2 with no reuse: 94 MB of L2 -> SM traffic for 5 MB of data at the
// Darcy shape, 13 / 29 us.  Here a block owns TM consecutive packed modes, stages X^ (and G^) and W of those modes in
// shared memory once, and register-tiles the small complex GEMMs:
//   forward   Y^[b, o]  = sum_i X^[b, i] W[i, o]                thread tile: up to 5 batch entries x 2 outputs
//   backward  G^x[b, i] = sum_o conj(W[i, o]) G^[b, o]          thread tile: up to 5 batch entries x 2 inputs
//             gW[i, o]  = sum_b conj(X^[b, i]) G^[b, o]         thread tile: up to 4 inputs x 4 outputs
//             glb[o]    = S * sum_b Re G^[b, o, mode 0]         (block of mode 0)
// Shared-memory pitches are padded so that the lanes of a warp (TM modes x 8 channel tiles) hit distinct banks.
// ----------------------------------------------------------------------------------------------
constexpr int kMixTileThreads = 256;
constexpr int kMixTileB = 5;       // batch entries per thread tile (register array size)

struct MixTileParams {
  const float2* xhat;    // [B][CI][M]
  const float2* ghat;    // backward: [B][CO][M]
  float2* yhat;          // forward:  [B][CO][M]
  float2* gxhat;         // backward: [B][CI][M], may be null
  const float2* w[4];    // weight blocks [CI][CO][Mblk]
  float2* gw[4];         // backward: may be null (gw[0] == nullptr)
  float* glb;            // backward: bias gradient from the DC mode, may be null
  float s_total;
  ModeDecode md;
  int B, CI, CO, TM;
  int bper;              // batch entries per thread tile (<= kMixTileB)
  int cit;               // input channels per thread tile of the gW role (<= 4)
};

struct MixTileSmem { int x, g, w, chp, wip, total; };   // float2 offsets; channel pitch; ci pitch of W
__host__ __device__ inline MixTileSmem mix_tile_smem(int B, int CI, int CO, int TM, bool backward) {
  MixTileSmem m;
  m.chp = TM + 1;                                  // one complex of padding per (., channel) row
  m.wip = CO * m.chp + 2;                          // ... and two per input-channel slab of W
  m.x = 0;
  int off = B * CI * m.chp;
  m.g = off;
  if (backward) off += B * CO * m.chp;
  m.w = off;
  off += CI * m.wip;
  m.total = off;
  return m;
}

// Staging loops keep eight global loads in flight per thread: a plain load -> store loop costs one memory round trip
// per trip, which for these small tiles IS the kernel (first version: 22 us for the forward product).
constexpr int kMixStageBatch = 8;

// stage [rows][TM] complex of a (B, C, M) spectrum tensor: rows = B * C, channel pitch chp
__device__ __forceinline__ void mix_tile_stage(float2* dst, const float2* src, int rows, int M, int pm0, int TM, int chp) {
  const int n = rows * TM;
  for (int idx0 = threadIdx.x; idx0 < n; idx0 += kMixStageBatch * kMixTileThreads) {
    float2 v[kMixStageBatch];
#pragma unroll
    for (int u = 0; u < kMixStageBatch; ++u) {
      const int idx = min(idx0 + u * kMixTileThreads, n - 1);
      const int r = idx / TM, t = idx - r * TM;
      v[u] = ldg_ordered(src + (long long)r * M + min(pm0 + t, M - 1));
    }
    DENO_SCHED_FENCE();
#pragma unroll
    for (int u = 0; u < kMixStageBatch; ++u) {
      const int idx = idx0 + u * kMixTileThreads;
      if (idx < n) {
        const int r = idx / TM, t = idx - r * TM;
        dst[r * chp + t] = (pm0 + t < M) ? v[u] : make_float2(0.f, 0.f);
      }
    }
  }
}

// stage the TM modes' weights [CI][CO][TM]; the (block, offset) of each mode is decoded once per block (woff, smem)
__device__ __forceinline__ void mix_tile_stage_w(const MixTileParams& p, float2* ws, const MixTileSmem& so, int pm0, const int* woff) {
  const int TM = p.TM, n = p.CI * p.CO * TM;
  for (int idx0 = threadIdx.x; idx0 < n; idx0 += kMixStageBatch * kMixTileThreads) {
    float2 v[kMixStageBatch];
#pragma unroll
    for (int u = 0; u < kMixStageBatch; ++u) {
      const int idx = min(idx0 + u * kMixTileThreads, n - 1);
      const int r = idx / TM, t = idx - r * TM;            // r = ci * CO + co
      const int wo = woff[t];                              // (blk << 28) | off
      const int blk = wo >> 28;
      const float2* wb = blk == 0 ? p.w[0] : (blk == 1 ? p.w[1] : (blk == 2 ? p.w[2] : p.w[3]));
      v[u] = ldg_ordered(wb + (long long)r * p.md.Mblk + (wo & 0x0FFFFFFF));
    }
    DENO_SCHED_FENCE();
#pragma unroll
    for (int u = 0; u < kMixStageBatch; ++u) {
      const int idx = idx0 + u * kMixTileThreads;
      if (idx < n) {
        const int r = idx / TM, t = idx - r * TM;
        const int co = r % p.CO, ci = r / p.CO;
        ws[ci * so.wip + co * so.chp + t] = (pm0 + t < p.md.M) ? v[u] : make_float2(0.f, 0.f);
      }
    }
  }
}

// (block, offset) of the block's modes, decoded once: woff[t] = (blk << 28) | off
__device__ __forceinline__ void mix_tile_decode(const MixTileParams& p, int pm0, int* woff) {
  if ((int)threadIdx.x < p.TM) {
    int blk, off;
    decode_mode(p.md, min(pm0 + (int)threadIdx.x, p.md.M - 1), blk, off);
    woff[threadIdx.x] = (blk << 28) | off;
  }
  __syncthreads();
}

// out[b][r][t] = sum_c A[b][c][t] * (conj?) W(c, r)[t] for the block's modes; W(c, r) = ws[c * cstride + r * rstride].
// BPER (batch entries per thread tile) is a template argument and every address is a base pointer plus c * stride, so
// the inner loop is loads + FMAs only (the first version spent 60 % of its instructions on index arithmetic and guards).
template <int BPER>
__device__ __forceinline__ void mix_tile_product(const float2* as, const float2* ws, float2* out, int B, int C, int R, int M,
                                                 int pm0, int TM, int chp, int w_cstride, int w_rstride, bool conj_w) {
  const int t = threadIdx.x % TM;
  const int nr2 = (R + 1) / 2;
  const int nbg = (B + BPER - 1) / BPER;
  const float sgn = conj_w ? -1.0f : 1.0f;
  for (int item = threadIdx.x / TM; item < nbg * nr2; item += kMixTileThreads / TM) {
    const int r2 = item % nr2, bg = item / nr2;
    const int r0 = 2 * r2, r1 = min(2 * r2 + 1, R - 1);
    const int b0 = bg * BPER;
    float ar[BPER][2], ai[BPER][2];
    const float2* ap[BPER];
#pragma unroll
    for (int j = 0; j < BPER; ++j) {
      ar[j][0] = ai[j][0] = ar[j][1] = ai[j][1] = 0.f;
      ap[j] = as + (min(b0 + j, B - 1) * C) * chp + t;
    }
    const float2* w0p = ws + r0 * w_rstride + t;
    const float2* w1p = ws + r1 * w_rstride + t;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float2 w0 = w0p[c * w_cstride], w1 = w1p[c * w_cstride];
      const float w0i = w0.y * sgn, w1i = w1.y * sgn;
#pragma unroll
      for (int j = 0; j < BPER; ++j) {
        const float2 a = ap[j][c * chp];
        ar[j][0] = fmaf(a.x, w0.x, fmaf(-a.y, w0i, ar[j][0]));
        ai[j][0] = fmaf(a.x, w0i, fmaf(a.y, w0.x, ai[j][0]));
        ar[j][1] = fmaf(a.x, w1.x, fmaf(-a.y, w1i, ar[j][1]));
        ai[j][1] = fmaf(a.x, w1i, fmaf(a.y, w1.x, ai[j][1]));
      }
    }
    if (pm0 + t < M) {
#pragma unroll
      for (int j = 0; j < BPER; ++j) {
        if (b0 + j < B) {
          out[((long long)(b0 + j) * R + r0) * M + pm0 + t] = make_float2(ar[j][0], ai[j][0]);
          if (2 * r2 + 1 < R) out[((long long)(b0 + j) * R + r0 + 1) * M + pm0 + t] = make_float2(ar[j][1], ai[j][1]);
        }
      }
    }
  }
}

__device__ __forceinline__ void mix_tile_product_any(int bper, const float2* as, const float2* ws, float2* out, int B, int C, int R,
                                                     int M, int pm0, int TM, int chp, int w_cstride, int w_rstride, bool conj_w) {
  switch (bper) {
    case 1: mix_tile_product<1>(as, ws, out, B, C, R, M, pm0, TM, chp, w_cstride, w_rstride, conj_w); break;
    case 2: mix_tile_product<2>(as, ws, out, B, C, R, M, pm0, TM, chp, w_cstride, w_rstride, conj_w); break;
    case 3: mix_tile_product<3>(as, ws, out, B, C, R, M, pm0, TM, chp, w_cstride, w_rstride, conj_w); break;
    case 4: mix_tile_product<4>(as, ws, out, B, C, R, M, pm0, TM, chp, w_cstride, w_rstride, conj_w); break;
    default: mix_tile_product<5>(as, ws, out, B, C, R, M, pm0, TM, chp, w_cstride, w_rstride, conj_w); break;
  }
}

// gW[i, o] = sum_b conj(X^[b, i]) G^[b, o] for the block's modes; thread tile CIT inputs x 4 outputs
template <int CIT>
__device__ __forceinline__ void mix_tile_wgrad(const MixTileParams& p, const float2* xs, const float2* gs, int chp, int pm0,
                                               const int* woff) {
  const int TM = p.TM, M = p.md.M;
  const int t = threadIdx.x % TM;
  const int nci = (p.CI + CIT - 1) / CIT, nco = (p.CO + 3) / 4;
  const int xstep = p.CI * chp, gstep = p.CO * chp;
  for (int item = threadIdx.x / TM; item < nci * nco; item += kMixTileThreads / TM) {
    const int oq = item % nco, iq = item / nco;
    const int i0 = iq * CIT, o0 = oq * 4;
    float ar[CIT][4], ai[CIT][4];
    const float2* xp[CIT];
    const float2* gp[4];
#pragma unroll
    for (int a = 0; a < CIT; ++a) {
      xp[a] = xs + min(i0 + a, p.CI - 1) * chp + t;
#pragma unroll
      for (int c = 0; c < 4; ++c) ar[a][c] = ai[a][c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) gp[c] = gs + min(o0 + c, p.CO - 1) * chp + t;
#pragma unroll 2
    for (int b = 0; b < p.B; ++b) {
      float2 xv[CIT], gv[4];
#pragma unroll
      for (int a = 0; a < CIT; ++a) xv[a] = xp[a][b * xstep];
#pragma unroll
      for (int c = 0; c < 4; ++c) gv[c] = gp[c][b * gstep];
#pragma unroll
      for (int a = 0; a < CIT; ++a) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // conj(x) * g = (xr gr + xi gi) + i (xr gi - xi gr)
          ar[a][c] = fmaf(xv[a].x, gv[c].x, fmaf(xv[a].y, gv[c].y, ar[a][c]));
          ai[a][c] = fmaf(xv[a].x, gv[c].y, fmaf(-xv[a].y, gv[c].x, ai[a][c]));
        }
      }
    }
    if (pm0 + t < M) {
      const int blk = woff[t] >> 28, off = woff[t] & 0x0FFFFFFF;
      float2* gb = blk == 0 ? p.gw[0] : (blk == 1 ? p.gw[1] : (blk == 2 ? p.gw[2] : p.gw[3]));
#pragma unroll
      for (int a = 0; a < CIT; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (i0 + a < p.CI && o0 + c < p.CO)
            gb[((long long)(i0 + a) * p.CO + o0 + c) * p.md.Mblk + off] = make_float2(ar[a][c], ai[a][c]);
    }
  }
}

__global__ void __launch_bounds__(kMixTileThreads) mode_mix_tile_kernel(MixTileParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float2* sm = reinterpret_cast<float2*>(smem_raw);
  const MixTileSmem so = mix_tile_smem(p.B, p.CI, p.CO, p.TM, false);
  const int pm0 = blockIdx.x * p.TM;
  int* woff = reinterpret_cast<int*>(sm + so.total);
  mix_tile_decode(p, pm0, woff);
  mix_tile_stage(sm + so.x, p.xhat, p.B * p.CI, p.md.M, pm0, p.TM, so.chp);
  mix_tile_stage_w(p, sm + so.w, so, pm0, woff);
  __syncthreads();
  // contraction over ci: W(ci, co) = ws[ci * wip + co * chp]
  mix_tile_product_any(p.bper, sm + so.x, sm + so.w, p.yhat, p.B, p.CI, p.CO, p.md.M, pm0, p.TM, so.chp, so.wip, so.chp, false);
}

__global__ void __launch_bounds__(kMixTileThreads) mode_mix_bwd_tile_kernel(MixTileParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float2* sm = reinterpret_cast<float2*>(smem_raw);
  const MixTileSmem so = mix_tile_smem(p.B, p.CI, p.CO, p.TM, true);
  const int pm0 = blockIdx.x * p.TM, TM = p.TM, M = p.md.M;
  float2* xs = sm + so.x;
  float2* gs = sm + so.g;
  float2* ws = sm + so.w;
  int* woff = reinterpret_cast<int*>(sm + so.total);
  mix_tile_decode(p, pm0, woff);
  if (p.gw[0] != nullptr) mix_tile_stage(xs, p.xhat, p.B * p.CI, M, pm0, TM, so.chp);
  mix_tile_stage(gs, p.ghat, p.B * p.CO, M, pm0, TM, so.chp);
  if (p.gxhat != nullptr) mix_tile_stage_w(p, ws, so, pm0, woff);
  __syncthreads();
  // G^x: contraction over co of conj(W(ci, co)): W as seen from (c = co, r = ci) is ws[r * wip + c * chp]
  if (p.gxhat != nullptr)
    mix_tile_product_any(p.bper, gs, ws, p.gxhat, p.B, p.CO, p.CI, M, pm0, TM, so.chp, so.chp, so.wip, true);
  // gW[i, o] = sum_b conj(X^[b, i]) G^[b, o]
  if (p.gw[0] != nullptr) {
    if (p.cit >= 4) mix_tile_wgrad<4>(p, xs, gs, so.chp, pm0, woff);
    else if (p.cit == 2) mix_tile_wgrad<2>(p, xs, gs, so.chp, pm0, woff);
    else mix_tile_wgrad<1>(p, xs, gs, so.chp, pm0, woff);
  }
  // bias gradient of the bypass conv from the DC mode (packed mode 0 lives in block 0, column 0)
  if (p.glb != nullptr && blockIdx.x == 0) {
    for (int o = threadIdx.x; o < p.CO; o += kMixTileThreads) {
      float acc = 0.f;
      for (int b = 0; b < p.B; ++b) acc += gs[(b * p.CO + o) * so.chp].x;
      p.glb[o] = acc * p.s_total;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// plane_synthesis: y = act( scale * sum_{k,l} c_l Re(spec e^{+i theta}) + sum_i W[o,i] u[i] + bias[o] )
// ----------------------------------------------------------------------------------------------
struct SynthParams {
  PlaneGeom g;           // geometry of the bypass input u (C = CI channels)
  int CO;                // produced channels
  long long sbo_out;     // bo stride of the produced tensors (sbi / sc equal the input's)
  const float2* spec;    // [NB'][CO][KH][MW]
  const float* u;
  const float* lw;       // bypass weight, W[o,i] = lw[o*lw_so + i*lw_si]
  long long lw_so, lw_si;
  const float* bias;     // may be null
  float* y;
  float* z;              // may be null; receives act'(pre-activation) for the backward pass
  const float2* twW;
  const float2* twH;
  const float* cl;
  float scale;
  int herm;
  int act;
  int R;                 // rows per tile (1 when H == 1)
  int PTS;               // points per tile (R*W when H > 1)
  int tiles;             // tiles per plane
};

struct SynthSmem {
  int xs, wls, bias, us, total;  // floats
  int PTSp, COp, MWp;
};

__host__ __device__ inline SynthSmem synth_smem(const SynthParams& p) {
  SynthSmem s;
  s.PTSp = round_up(p.PTS, 4);
  s.COp = round_up(p.CO, 8);
  s.MWp = round_up(p.g.MW, 4);
  int off = 0;
  s.xs = off;   off += p.g.C * s.PTSp;
  s.wls = off;  off += p.g.C * s.COp;
  s.bias = off; off += s.COp;
  s.us = off;   off += s.COp * p.R * s.MWp * 2;
  s.total = off;
  return s;
}

template <int ACT>
__global__ void plane_synthesis_kernel(SynthParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int H = p.g.H, W = p.g.W, KH = p.g.KH, MW = p.g.MW, CI = p.g.C, CO = p.CO;
  const int KH4 = round_up(KH, 4);
  const SynthSmem so = synth_smem(p);
  const int PTSp = so.PTSp, COp = so.COp, MWp = so.MWp, R = p.R;
  float* xs = smem + so.xs;
  float* wls = smem + so.wls;
  float* bs = smem + so.bias;
  float2* Us = reinterpret_cast<float2*>(smem + so.us);

  const int bp = blockIdx.x / p.tiles;
  const int tile = blockIdx.x - bp * p.tiles;
  const int p0 = tile * p.PTS;
  const int npts = min(p.PTS, H * W - p0);
  const int x0 = p0 / W;
  const int trows = (H > 1) ? min(R, H - x0) : 1;

  // bypass operands
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int i = warp; i < CI; i += nwarps)
    for (int o = lane; o < COp; o += 32)
      wls[i * COp + o] = (o < CO) ? p.lw[(long long)o * p.lw_so + (long long)i * p.lw_si] : 0.f;
  for (int o = tid; o < COp; o += nt) bs[o] = (p.bias != nullptr && o < CO) ? p.bias[o] : 0.f;
  for (int i = warp; i < CI; i += nwarps) {
    const float* src = p.u + plane_base(p.g, bp, i) + p0;
    for (int pt = lane; pt < npts; pt += 32) xs[i * PTSp + pt] = src[pt];
  }
  // stage A: U[o][rr][l] = f_l * sum_k spec[o][k][l] e^{+i phi(k, x0+rr)}
  for (int task = tid; task < COp * MWp; task += nt) {
    const int o = task / MWp, l = task - o * MWp;
    float ar[kSynthMaxRows], ai[kSynthMaxRows];
#pragma unroll
    for (int rr = 0; rr < kSynthMaxRows; ++rr) ar[rr] = ai[rr] = 0.f;
    float f = 0.f;
    if (o < CO && l < MW) {
      f = p.scale * (p.herm ? p.cl[l] : 1.0f);
      const float2* sp = p.spec + ((long long)bp * CO + o) * KH * MW + l;
      for (int k = 0; k < KH; ++k) {
        const float2 s = sp[k * MW];
#pragma unroll
        for (int rr = 0; rr < kSynthMaxRows; ++rr) {
          if (rr < trows) {
            const float2 e = p.twH[(x0 + rr) * KH4 + k];
            ar[rr] = fmaf(s.x, e.x, fmaf(-s.y, e.y, ar[rr]));
            ai[rr] = fmaf(s.x, e.y, fmaf(s.y, e.x, ai[rr]));
          }
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < kSynthMaxRows; ++rr)
      if (rr < R) Us[(o * R + rr) * MWp + l] = make_float2(ar[rr] * f, ai[rr] * f);
  }
  __syncthreads();

  // stage B: one point per thread, eight produced channels at a time
  const int half = MWp / 2;
  for (int pt = tid; pt < npts; pt += nt) {
    const int P = p0 + pt;
    const int xr = P / W;
    const int yv = P - xr * W;
    const int rr = xr - x0;
    const float4* twp = reinterpret_cast<const float4*>(p.twW + (long long)yv * MWp);
    const long long obase = (long long)(bp / p.g.NBI) * p.sbo_out + (long long)(bp % p.g.NBI) * p.g.sbi + P;
    for (int og = 0; og < COp; og += 8) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bs[og + j];
      for (int i = 0; i < CI; ++i) {
        const float xv = xs[i * PTSp + pt];
        const float4* w4 = reinterpret_cast<const float4*>(wls + i * COp + og);
        const float4 wa = w4[0], wb = w4[1];
        acc[0] = fmaf(xv, wa.x, acc[0]); acc[1] = fmaf(xv, wa.y, acc[1]);
        acc[2] = fmaf(xv, wa.z, acc[2]); acc[3] = fmaf(xv, wa.w, acc[3]);
        acc[4] = fmaf(xv, wb.x, acc[4]); acc[5] = fmaf(xv, wb.y, acc[5]);
        acc[6] = fmaf(xv, wb.z, acc[6]); acc[7] = fmaf(xv, wb.w, acc[7]);
      }
      for (int lc = 0; lc < half; ++lc) {
        const float4 t = twp[lc];  // (c0, s0, c1, s1)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 uu = reinterpret_cast<const float4*>(Us + ((og + j) * R + rr) * MWp)[lc];
          acc[j] = fmaf(uu.x, t.x, fmaf(-uu.y, t.y, fmaf(uu.z, t.z, fmaf(-uu.w, t.w, acc[j]))));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = og + j;
        if (o < CO) {
          const long long a = obase + (long long)o * p.g.sc;
          if (p.z != nullptr) {
            float yv, dv;
            act_apply_grad(ACT, acc[j], yv, dv);
            p.z[a] = dv;
            p.y[a] = yv;
          } else {
            p.y[a] = act_apply(ACT, acc[j]);
          }
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------
// bypass_wgrad: partial[cta][o*CI+i] = sum_pt gz[o,pt] x[i,pt];  partial[cta][CO*CI+o] = sum_pt gz[o,pt]
// ----------------------------------------------------------------------------------------------
struct WgradParams {
  PlaneGeom gx;        // geometry of x (C = CI)
  int CO;
  long long sbo_g;     // bo stride of gz
  const float* gz;
  const float* x;
  float* partial;      // [ncta][CO*CI + CO]
  int PTS;             // points per tile (multiple of 4)
  int tiles;           // tiles per plane
};

// shared memory: gz tile [CO][pitch], x tile [CIp + 1][pitch] (last row = zeros), pitch = PTS + 4 with
// PTS a multiple of 32 so that eight consecutive rows start in eight different bank quads.
__host__ __device__ inline int wgrad_smem_floats(const WgradParams& p) {
  return (p.CO + round_up(p.gx.C, 8) + 1) * (p.PTS + 4);
}

__global__ void bypass_wgrad_partial_kernel(WgradParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  const int CI = p.gx.C, CO = p.CO, CIp = round_up(CI, 8), JB = (CIp + 31) / 32;
  const int PTS = p.PTS, pitch = PTS + 4;
  float* gs = smem;                 // [CO][pitch]
  float* xs = smem + CO * pitch;    // [CIp + 1][pitch]
  const int bp = blockIdx.x / p.tiles;
  const int tile = blockIdx.x - bp * p.tiles;
  const int p0 = tile * PTS;
  const int npts = min(PTS, p.gx.H * p.gx.W - p0);
  const long long gbase = (long long)(bp / p.gx.NBI) * p.sbo_g + (long long)(bp % p.gx.NBI) * p.gx.sbi + p0;

  for (int o = warp; o < CO; o += nwarps) {
    const float* src = p.gz + gbase + (long long)o * p.gx.sc;
    for (int pt = lane; pt < PTS; pt += 32) gs[o * pitch + pt] = (pt < npts) ? src[pt] : 0.f;
  }
  for (int i = warp; i <= CIp; i += nwarps) {
    const float* src = p.x + plane_base(p.gx, bp, i < CI ? i : 0) + p0;
    for (int pt = lane; pt < PTS; pt += 32) xs[i * pitch + pt] = (pt < npts && i < CI) ? src[pt] : 0.f;
  }
  __syncthreads();
  float* out = p.partial + (long long)blockIdx.x * (CO * CI + CO);
  // task = (o, block of 32 input channels, iq); the thread owns input channels iq, iq+8, iq+16, iq+24 of the block
  for (int task = tid; task < CO * JB * 8; task += nt) {
    const int iq = task & 7;
    const int jb = (task >> 3) % JB;
    const int o = (task >> 3) / JB;
    int irow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = jb * 32 + iq + 8 * j;
      irow[j] = i < CIp ? i : CIp;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float4* g4 = reinterpret_cast<const float4*>(gs + o * pitch);
    const float4* x0 = reinterpret_cast<const float4*>(xs + irow[0] * pitch);
    const float4* x1 = reinterpret_cast<const float4*>(xs + irow[1] * pitch);
    const float4* x2 = reinterpret_cast<const float4*>(xs + irow[2] * pitch);
    const float4* x3 = reinterpret_cast<const float4*>(xs + irow[3] * pitch);
#pragma unroll 2
    for (int v = 0; v < PTS / 4; ++v) {
      const float4 g = g4[v];
      float4 t = x0[v];
      a0 = fmaf(g.x, t.x, fmaf(g.y, t.y, fmaf(g.z, t.z, fmaf(g.w, t.w, a0))));
      t = x1[v];
      a1 = fmaf(g.x, t.x, fmaf(g.y, t.y, fmaf(g.z, t.z, fmaf(g.w, t.w, a1))));
      t = x2[v];
      a2 = fmaf(g.x, t.x, fmaf(g.y, t.y, fmaf(g.z, t.z, fmaf(g.w, t.w, a2))));
      t = x3[v];
      a3 = fmaf(g.x, t.x, fmaf(g.y, t.y, fmaf(g.z, t.z, fmaf(g.w, t.w, a3))));
    }
    const int i = jb * 32 + iq;
    if (i < CI) out[o * CI + i] = a0;
    if (i + 8 < CI) out[o * CI + i + 8] = a1;
    if (i + 16 < CI) out[o * CI + i + 16] = a2;
    if (i + 24 < CI) out[o * CI + i + 24] = a3;
  }
  for (int o = tid; o < CO; o += nt) {   // bias gradient (fixed summation order)
    const float4* g4 = reinterpret_cast<const float4*>(gs + o * pitch);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int v = 0; v < PTS / 4; ++v) {
      const float4 g = g4[v];
      s0 += g.x; s1 += g.y; s2 += g.z; s3 += g.w;
    }
    out[CO * CI + o] = (s0 + s1) + (s2 + s3);
  }
}

struct WgradReduceParams {
  const float* partial;
  int ncta, nj, nw;   // nj = CO*CI + CO outputs, the first nw go to glw, the rest to glb
  int pitch;          // floats between consecutive partial rows (>= nj)
  float* glw;         // may be null
  float* glb;         // may be null
};

// blockDim = (32 outputs, kReduceSlices slices); grid = ceil(nj / 32).  Fixed summation order -> deterministic.
// 32 slices x 8 loads in flight: the ~600 partial rows of a layer's weight gradient are three memory round trips
// (with 8 slices it was ten, ~10 us per launch, eight launches per training step).
constexpr int kReduceSlices = 32;
__global__ void __launch_bounds__(32 * kReduceSlices) bypass_wgrad_reduce_kernel(WgradReduceParams p) {
  DENO_PDL_SYNC();
  DENO_DYN_SMEM(smem_raw);
  float* red = reinterpret_cast<float*>(smem_raw);  // [kReduceSlices][32]
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int slice = threadIdx.y;
  float s = 0.f;
  if (j < p.nj) {
    for (int c0 = slice; c0 < p.ncta; c0 += 8 * kReduceSlices) {   // eight loads in flight per thread
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + kReduceSlices * u;
        v[u] = ldg_ordered(p.partial + (long long)min(c, p.ncta - 1) * p.pitch + j);
        if (c >= p.ncta) v[u] = 0.f;
      }
      DENO_SCHED_FENCE();
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
  }
  red[slice * 32 + threadIdx.x] = s;
  __syncthreads();
  if (slice == 0 && j < p.nj) {
    float t = 0.f;
    for (int k = 0; k < kReduceSlices; ++k) t += red[k * 32 + threadIdx.x];
    if (j < p.nw) {
      if (p.glw != nullptr) p.glw[j] = t;
    } else if (p.glb != nullptr) {
      p.glb[j - p.nw] = t;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// adam_flat: one fused Adam step over a flat fp32 parameter buffer (complex64 weights as their real
// view, which is also how torch.optim.Adam treats them).  Same update as torch.optim.Adam(amsgrad=False):
//   g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The step count t is read from device memory so the launch can be replayed from a CUDA graph.
// ----------------------------------------------------------------------------------------------
struct AdamParams {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  float lr, beta1, beta2, eps, weight_decay;
  const float* step;   // device scalar: t (already incremented for this step)
};

__global__ void adam_flat_kernel(AdamParams a) {
  DENO_PDL_SYNC();
  const float t = a.step[0];
  const float bc1 = 1.0f - powf(a.beta1, t);
  const float bc2 = 1.0f - powf(a.beta2, t);
  const float step_size = a.lr / bc1;
  const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = a.p[i];
    const float gv = fmaf(a.weight_decay, pv, a.g[i]);
    const float mv = fmaf(a.beta1, a.m[i], (1.0f - a.beta1) * gv);
    const float vv = fmaf(a.beta2, a.v[i], (1.0f - a.beta2) * gv * gv);
    a.m[i] = mv;
    a.v[i] = vv;
    a.p[i] = pv - step_size * (mv / fmaf(sqrtf(vv), inv_sqrt_bc2, a.eps));
  }
}

}  // namespace deno
