// HBM-bound glue around the epi / temporal attention: GroupNorm + token layout, LayerNorm,
// GEGLU, residual updates and the inverse layout.  All kernels are coalesced on both the
// read and the write side (layout changes go through a padded shared-memory tile), use
// 8/16-byte vector accesses where the shape allows and warp-shuffle reductions.
#include <cooperative_groups.h>
#include <type_traits>

#include "common.cuh"

namespace cvd {

// token index of (b, f, p) in the two layouts the reference uses
__device__ __forceinline__ size_t token_of(int b, int f, int p, int F, int HW, int layout) {
  return layout == CVD_LAYOUT_BF_HW ? ((size_t)b * F + f) * HW + p : ((size_t)b * HW + p) * F + f;
}

// ---------------------------------------------------------------------------------------
// GroupNorm statistics: one CTA per (b, f, group).  resnet.py:30-40 applies nn.GroupNorm
// per frame, so a group is cpg channels x HW pixels of ONE frame (channel stride F*HW).
// ---------------------------------------------------------------------------------------
// 16-byte vector loads of the native dtype, widened to float
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* v) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

// Single pass, shifted by the group's first element so that sum((x-x0)^2) - n*mean(x-x0)^2 stays
// well conditioned; 16-byte loads when HW allows.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const T* __restrict__ x, float* __restrict__ stats,
                                                              int C, int F, int HW, int groups, float eps) {
  const int g = blockIdx.x % groups;
  const int f = (blockIdx.x / groups) % F;
  const int b = blockIdx.x / (groups * F);
  const int cpg = C / groups;
  const int n = cpg * HW;
  const T* base = x + (((size_t)b * C + (size_t)g * cpg) * F + f) * HW;
  const size_t cstride = (size_t)F * HW;
  const float x0 = load_as_float(base, 0);
  float s = 0.f, q = 0.f;
  if (VEC) {
    constexpr int V = Vec16<T>::N;
    const int per_c = HW / V;
    for (int i = threadIdx.x; i < cpg * per_c; i += blockDim.x) {
      float v[V];
      Vec16<T>::load(base + (size_t)(i / per_c) * cstride + (size_t)(i % per_c) * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float d = v[k] - x0;
        s += d;
        q = fmaf(d, d, q);
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float d = load_as_float(base, (size_t)(i / HW) * cstride + (i % HW)) - x0;
      s += d;
      q = fmaf(d, d, q);
    }
  }
  __shared__ float rs[8], rq[8];
  s = warp_sum(s);
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rq[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      ts += rs[i];
      tq += rq[i];
    }
    const float md = ts / (float)n;
    const float var = fmaxf(tq / (float)n - md * md, 0.f);
    stats[2 * (size_t)blockIdx.x] = x0 + md;
    stats[2 * (size_t)blockIdx.x + 1] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------------------------------
// video [B,C,F,HW] -> bf16 tokens [T, C] through a 64-channel x 32-pixel tile.
//   MODE 0: y = groupnorm(x)                (epi_module.py:163-165, motion_module.py:377-378)
//   MODE 1: y = tok_in + x                  (pose feature add, attention_processor.py:325)
// ---------------------------------------------------------------------------------------
constexpr int kTileC = 64, kTileP = 32;

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
video_to_tokens_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const __nv_bfloat16* __restrict__ tok_in,
                       __nv_bfloat16* __restrict__ y, int C, int F, int HW, int groups, int layout) {
  __shared__ float tile[kTileC][kTileP + 1];
  const int p0 = blockIdx.x * kTileP, c0 = blockIdx.y * kTileC;
  const int f = blockIdx.z % F, b = blockIdx.z / F;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int cpg = C / groups;
#pragma unroll
  for (int i = 0; i < kTileC / 8; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) {
      v = load_as_float(x, (((size_t)b * C + c) * F + f) * HW + p);
      if (MODE == 0) {
        const float* st = stats + 2 * ((size_t)(b * F + f) * groups + c / cpg);
        v = (v - st[0]) * st[1] * gamma[c] + beta[c];
      }
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
  // write: 32 lanes cover 64 channels (2 each) of one token -> 128-byte rows
#pragma unroll
  for (int i = 0; i < kTileP / 8; ++i) {
    const int p = p0 + ty + 8 * i, c = c0 + 2 * tx;
    if (p < HW && c < C) {
      const size_t o = token_of(b, f, p, F, HW, layout) * C + c;
      float v0 = tile[2 * tx][ty + 8 * i], v1 = tile[2 * tx + 1][ty + 8 * i];
      if (MODE == 1) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(tok_in + o);
        v0 += __bfloat162float(t.x);
        v1 += __bfloat162float(t.y);
      }
      *reinterpret_cast<__nv_bfloat162*>(y + o) = __floats2bfloat162_rn(v0, v1);
    }
  }
}

// ---------------------------------------------------------------------------------------
// tokens bf16 [T, C] (+bias) -> video layout + residual   (epi_module.py:185-187)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
tokens_to_video_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ bias,
                       const T* __restrict__ residual, T* __restrict__ out, int C, int F, int HW, int layout) {
  __shared__ float tile[kTileP][kTileC + 1];
  const int p0 = blockIdx.x * kTileP, c0 = blockIdx.y * kTileC;
  const int f = blockIdx.z % F, b = blockIdx.z / F;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kTileP / 8; ++i) {
    const int p = p0 + ty + 8 * i, c = c0 + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (p < HW && c < C) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(tok + token_of(b, f, p, F, HW, layout) * C + c);
      v0 = __bfloat162float(t.x) + (bias ? bias[c] : 0.f);
      v1 = __bfloat162float(t.y) + (bias ? bias[c + 1] : 0.f);
    }
    tile[ty + 8 * i][2 * tx] = v0;
    tile[ty + 8 * i][2 * tx + 1] = v1;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kTileC / 8; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    if (c < C && p < HW) {
      const size_t o = (((size_t)b * C + c) * F + f) * HW + p;
      store_from_float(out, o, load_as_float(residual, o) + tile[tx][ty + 8 * i]);
    }
  }
}


// ---------------------------------------------------------------------------------------
// tokens -> video for the fp32 stream, full-row version: a CTA owns 32 pixels x ALL channels of one frame.  The
// token side is then one contiguous 32*C*2-byte block per CTA (BF_HW layout; 32 rows of C*2 bytes in the temporal
// layout) read with 16-byte loads -- the 64-channel tile kernel above touches every 640-byte token row from five
// different CTAs at five different times (DRAM page locality) -- and lands in shared memory as bf16 [32][C+2]
// (odd word pitch: the per-channel column reads of the video side are conflict-free).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tokens_to_video_rows_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ bias,
                            const float* __restrict__ residual, float* __restrict__ out, int C, int F, int HW, int layout) {
  extern __shared__ __align__(16) unsigned char t2v_smem[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(t2v_smem);      // [32][C/2 + 1] words (bf16 pairs)
  const int pitch = C / 2 + 1;
  const int nblk = HW / kTileP;
  const int pb = blockIdx.x % nblk;
  const int f = (blockIdx.x / nblk) % F;
  const int b = blockIdx.x / (nblk * F);
  const int p0 = pb * kTileP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c8 = C / 8;
  for (int i = threadIdx.x; i < kTileP * c8; i += 256) {
    const int pl = i / c8, v = i % c8;
    const uint4 t = *reinterpret_cast<const uint4*>(tok + token_of(b, f, p0 + pl, F, HW, layout) * C + v * 8);
    uint32_t* d = tile + pl * pitch + v * 4;
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  }
  __syncthreads();
  const size_t vbase = ((size_t)b * C * F + f) * HW + p0 + lane;      // + c * F * HW
  const size_t cstride = (size_t)F * HW;
  const __nv_bfloat16* th = reinterpret_cast<const __nv_bfloat16*>(tile) + (size_t)lane * pitch * 2;
#pragma unroll 8
  for (int c = warp; c < C; c += 8) {
    const size_t o = vbase + (size_t)c * cstride;
    out[o] = residual[o] + __bfloat162float(th[c]) + (bias ? bias[c] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------
// LayerNorm: one warp per token, the row lives in registers (C <= 1280), two-pass variance.
// ---------------------------------------------------------------------------------------
template <int PAIRS>   // float2 pairs per lane = C / 64
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ pe, __nv_bfloat16* __restrict__ y, int tokens, int seq_len, float eps) {
  constexpr int C = PAIRS * 64;
  const int lane = threadIdx.x & 31;
  const size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= (size_t)tokens) return;
  const float2* row = reinterpret_cast<const float2*>(x + t * C);
  float2 v[PAIRS];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    v[j] = row[lane + 32 * j];
    s += v[j].x + v[j].y;
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    const float dx = v[j].x - mean, dy = v[j].y - mean;
    q = fmaf(dx, dx, fmaf(dy, dy, q));
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  const float2* g2 = reinterpret_cast<const float2*>(gamma);
  const float2* b2 = reinterpret_cast<const float2*>(beta);
  const float2* p2 = pe ? reinterpret_cast<const float2*>(pe + (t % seq_len) * C) : nullptr;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(y + t * C);
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    const int i = lane + 32 * j;
    const float2 g = g2[i], b = b2[i];
    float ox = (v[j].x - mean) * rstd * g.x + b.x, oy = (v[j].y - mean) * rstd * g.y + b.y;
    if (p2) {
      const float2 p = p2[i];
      ox += p.x;
      oy += p.y;
    }
    out[i] = __floats2bfloat162_rn(ox, oy);
  }
}

// ---------------------------------------------------------------------------------------
// Fused residual update + LayerNorm: x += (y + bias) * scale (in place, fp32), n = LN(x) (+pe) -> bf16.
// One warp per token, the row stays in registers: 12 bytes per element cross HBM instead of the 16 of
// residual_add_kernel followed by layernorm_kernel.  y == nullptr: x += bias only (the deferred bias of
// proj_in, whose GEMM writes fp32 directly).
// ---------------------------------------------------------------------------------------
template <int PAIRS>
__global__ void __launch_bounds__(256)
residual_layernorm_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ yin, const float* __restrict__ bias,
                          float scale, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ pe, __nv_bfloat16* __restrict__ n_out, int tokens, int seq_len, float eps) {
  constexpr int C = PAIRS * 64;
  const int lane = threadIdx.x & 31;
  const size_t t = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= (size_t)tokens) return;
  float2* row = reinterpret_cast<float2*>(x + t * C);
  const __nv_bfloat162* yrow = yin ? reinterpret_cast<const __nv_bfloat162*>(yin + t * C) : nullptr;
  const float2* bi = bias ? reinterpret_cast<const float2*>(bias) : nullptr;
  float2 v[PAIRS];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    const int i = lane + 32 * j;
    v[j] = row[i];
    float2 d = make_float2(0.f, 0.f);
    if (yrow) d = __bfloat1622float2(yrow[i]);
    if (bi) {
      const float2 b = bi[i];
      d.x += b.x;
      d.y += b.y;
    }
    v[j].x = fmaf(d.x, scale, v[j].x);
    v[j].y = fmaf(d.y, scale, v[j].y);
    row[i] = v[j];
    s += v[j].x + v[j].y;
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    const float dx = v[j].x - mean, dy = v[j].y - mean;
    q = fmaf(dx, dx, fmaf(dy, dy, q));
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  const float2* g2 = reinterpret_cast<const float2*>(gamma);
  const float2* b2 = reinterpret_cast<const float2*>(beta);
  const float2* p2 = pe ? reinterpret_cast<const float2*>(pe + (t % seq_len) * C) : nullptr;
  __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(n_out + t * C);
#pragma unroll
  for (int j = 0; j < PAIRS; ++j) {
    const int i = lane + 32 * j;
    const float2 g = g2[i], b = b2[i];
    float ox = (v[j].x - mean) * rstd * g.x + b.x, oy = (v[j].y - mean) * rstd * g.y + b.y;
    if (p2) {
      const float2 p = p2[i];
      ox += p.x;
      oy += p.y;
    }
    out[i] = __floats2bfloat162_rn(ox, oy);
  }
}

// ---------------------------------------------------------------------------------------
// GEGLU: out = h * gelu_erf(g), 8 bf16 (16 bytes) per thread.
// ---------------------------------------------------------------------------------------
// Exact-form (erf) GELU of diffusers' GEGLU (F.gelu default).  erff() costs ~30 instructions and makes the kernel
// compute-bound (65 % of the HBM roofline); here erfc(|z|) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),
// t = 1 / (1 + p |z|) (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7, i.e. 2^-15 of the bf16 output's last bit) with
// MUFU reciprocal / exp2 (rcp.approx / ex2.approx: ~1 ulp), and 1 + erf(z) is assembled without cancellation on the negative side.
__device__ __forceinline__ float gelu_erf(float g) {
  const float z = g * 0.70710678118654752f;
  const float az = fabsf(z);
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * az * az));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = poly * t * ex;   // erfc(|z|)
  return 0.5f * g * (z < 0.f ? e : 2.0f - e);
}

__global__ void __launch_bounds__(256) geglu_kernel(const uint4* __restrict__ hg, uint4* __restrict__ out,
                                                    size_t tokens, int inner8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * inner8) return;
  const size_t t = i / inner8, c = i % inner8;
  const uint4 hv = hg[t * 2 * inner8 + c], gv = hg[t * 2 * inner8 + inner8 + c];
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv);
  const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
  uint4 ov;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 h = __bfloat1622float2(h2[j]), g = __bfloat1622float2(g2[j]);
    o2[j] = __floats2bfloat162_rn(h.x * gelu_erf(g.x), h.y * gelu_erf(g.y));
  }
  out[i] = ov;
}

// ---------------------------------------------------------------------------------------
// x += (y + bias) * scale   (fp32 residual stream), optional bf16 copy; 4 elements per thread.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_add_kernel(float4* __restrict__ x, const uint2* __restrict__ y, const float4* __restrict__ bias, float scale,
                    uint2* __restrict__ xb, size_t n4, int C4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 xv = x[i];
  const uint2 yv = y[i];
  const float2 y0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.x));
  const float2 y1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv.y));
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b = bias[i % C4];
  xv.x += (y0.x + b.x) * scale;
  xv.y += (y0.y + b.y) * scale;
  xv.z += (y1.x + b.z) * scale;
  xv.w += (y1.y + b.w) * scale;
  x[i] = xv;
  if (xb) {
    uint2 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(xv.x, xv.y);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(xv.z, xv.w);
    xb[i] = o;
  }
}

// out = a * scale + b  (bf16, 8 per thread)   -- attention_processor.py:325
__global__ void __launch_bounds__(256) scale_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                        float scale, uint4* __restrict__ out, size_t n8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 av = a[i], bv = b[i];
  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&av);
  const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bv);
  uint4 ov;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = __bfloat1622float2(a2[j]), y = __bfloat1622float2(b2[j]);
    o2[j] = __floats2bfloat162_rn(fmaf(x.x, scale, y.x), fmaf(x.y, scale, y.y));
  }
  out[i] = ov;
}

// ---------------------------------------------------------------------------------------
// Fused GroupNorm + token layout for the fp32 residual stream: ONE read of x, ONE write of y.
// A thread-block CLUSTER owns kFusedCB = 80 channels (8 / 4 / 2 whole groups at C = 320 / 640 /
// 1280) of one frame; each of its CTAs takes 128 (or 64) pixels of them and keeps its 80 x 128
// slab IN REGISTERS (10 independent float4 loads per thread, all in flight at once).  Per-channel
// shifted sums meet in distributed shared memory (one cluster.sync), every CTA derives the group
// statistics from all ranks' partials, normalises its registers, transposes them through a
// conflict-free shared tile and writes 160-byte token segments.  The two-kernel version above
// streams x twice and runs 10 240 eight-kilobyte CTAs for the layout pass.
// ---------------------------------------------------------------------------------------
constexpr int kFusedCB = 80;

template <int PPL>   // pixels per lane: 4 (128 pixels per CTA, float4) or 2 (64 pixels, float2)
__global__ void __launch_bounds__(256)
groupnorm_tokens_fused_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              __nv_bfloat16* __restrict__ y, int C, int F, int HW, int groups, float eps, int layout,
                              int csize) {
  namespace cg = cooperative_groups;
  constexpr int CB = kFusedCB, PX = 32 * PPL, PITCH = PX + 1, NCH = CB / 8;
  extern __shared__ float gn_smem[];
  float* tile = gn_smem;                       // [CB][PITCH], column of pixel p: (p % PPL) * 32 + p / PPL
  float* psum = tile + CB * PITCH;             // [CB] this CTA's partial sums (read by the whole cluster)
  float* psq = psum + CB;
  float* cscale = psq + CB;
  float* cshift = cscale + CB;
  float* gsum = cshift + CB;                   // [8] per-group partials of this CTA (read by the whole cluster)
  float* gsq = gsum + 8;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = blockIdx.x % csize;
  const int item = blockIdx.x / csize;
  const int nblk = C / CB;
  const int cb = item % nblk;
  const int f = (item / nblk) % F;
  const int b = item / (nblk * F);
  const int c0 = cb * CB;
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t cstride = (size_t)F * HW;
  const float* base = x + (((size_t)b * C + c0) * F + f) * HW;      // channel c0 + c lives at base + c * cstride
  const int p0 = rank * PX;

  float v[NCH][PPL];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const float* src = base + (size_t)(warp + 8 * i) * cstride + p0 + lane * PPL;
    if (PPL == 4) {
      const float4 t = *reinterpret_cast<const float4*>(src);
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][PPL - 1] = t.w;
    } else {
      const float2 t = *reinterpret_cast<const float2*>(src);
      v[i][0] = t.x; v[i][1] = t.y;
    }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = warp + 8 * i;
    const float x0 = __ldg(base + (size_t)(c / cpg * cpg) * cstride);     // shift: first element of the group (conditioning)
    float sm = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < PPL; ++k) {
      const float d = v[i][k] - x0;
      sm += d;
      q = fmaf(d, d, q);
    }
    sm = warp_sum(sm);
    q = warp_sum(q);
    if (lane == 0) {
      psum[c] = sm;
      psq[c] = q;
    }
  }
  __syncthreads();
  const int ngrp = CB / cpg;                   // whole groups in this channel block (<= 8)
  if (threadIdx.x < ngrp) {                    // this CTA's partial sums per GROUP: 2 * ngrp values cross the cluster
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < cpg; ++i) {
      ts += psum[threadIdx.x * cpg + i];
      tq += psq[threadIdx.x * cpg + i];
    }
    gsum[threadIdx.x] = ts;
    gsq[threadIdx.x] = tq;
  }
  if (csize > 1) cluster.sync(); else __syncthreads();
  if (threadIdx.x < CB) {
    const int c = threadIdx.x, g = c / cpg;
    float ts = 0.f, tq = 0.f;
    if (csize > 1) {
      for (int r = 0; r < csize; ++r) {
        ts += cluster.map_shared_rank(gsum, r)[g];
        tq += cluster.map_shared_rank(gsq, r)[g];
      }
    } else {
      ts = gsum[g];
      tq = gsq[g];
    }
    const float n = (float)cpg * (float)HW;
    const float x0 = __ldg(base + (size_t)(g * cpg) * cstride);
    const float md = ts / n;
    const float rstd = rsqrtf(fmaxf(tq / n - md * md, 0.f) + eps);
    const float sc = rstd * gamma[c0 + c];
    cscale[c] = sc;
    cshift[c] = beta[c0 + c] - (x0 + md) * sc;
  }
  __syncthreads();
  // This CTA has read everything it needs from its peers (the values are in registers / local shared memory): signal
  // it now, relaxed (nothing is published), and collect the peers' signals only before exiting -- a CTA must not exit
  // while a peer can still read its partial sums, but nobody has to WAIT for that in the middle of the kernel.
  if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = warp + 8 * i;
    const float sc = cscale[c], sh = cshift[c];
#pragma unroll
    for (int k = 0; k < PPL; ++k) tile[c * PITCH + k * 32 + lane] = fmaf(v[i][k], sc, sh);
  }
  __syncthreads();
#pragma unroll 4
  for (int e = threadIdx.x; e < PX * (CB / 2); e += 256) {
    const int pl = e / (CB / 2), c2 = e % (CB / 2);
    const int col = (pl % PPL) * 32 + pl / PPL;
    const size_t o = token_of(b, f, p0 + pl, F, HW, layout) * C + c0 + 2 * c2;
    *reinterpret_cast<__nv_bfloat162*>(y + o) =
        __floats2bfloat162_rn(tile[(2 * c2) * PITCH + col], tile[(2 * c2 + 1) * PITCH + col]);
  }
  if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

template <int PPL>
static int launch_groupnorm_fused(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y, int B, int C,
                                  int F, int HW, int groups, float eps, int layout, cudaStream_t st) {
  constexpr int PX = 32 * PPL;
  const int csize = HW / PX;
  const size_t smem = (size_t)(kFusedCB * (PX + 1) + 4 * kFusedCB + 16) * sizeof(float);
  CVD_CUDA(ensure_dynamic_smem((const void*)groupnorm_tokens_fused_kernel<PPL>, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * F * (C / kFusedCB) * csize));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CVD_CUDA(cudaLaunchKernelEx(&cfg, groupnorm_tokens_fused_kernel<PPL>, x, gamma, beta, y, C, F, HW, groups, eps, layout, csize));
  CVD_LAUNCH_CHECK("groupnorm_tokens_fused_kernel");
  return CVD_OK;
}

template <typename T>
static int launch_groupnorm_tokens(const void* x, const float* gamma, const float* beta, void* y, float* stats, int B,
                                   int C, int F, int HW, int groups, float eps, int layout, cudaStream_t st) {
  const bool vec = (HW % Vec16<T>::N == 0) && (((uintptr_t)x & 15) == 0);
  const int cpg = C / groups;
  if (std::is_same<T, float>::value && vec && C % kFusedCB == 0 && kFusedCB % cpg == 0) {
    // single pass over HBM, slab in registers (stats stays untouched: it is scratch of the two-kernel path)
    const int cs4 = HW / 128, cs2 = HW / 64;
    if (HW % 128 == 0 && (cs4 == 1 || cs4 == 2 || cs4 == 4 || cs4 == 8))
      return launch_groupnorm_fused<4>((const float*)x, gamma, beta, (__nv_bfloat16*)y, B, C, F, HW, groups, eps, layout, st);
    if (HW % 64 == 0 && (cs2 == 1 || cs2 == 2 || cs2 == 4 || cs2 == 8))
      return launch_groupnorm_fused<2>((const float*)x, gamma, beta, (__nv_bfloat16*)y, B, C, F, HW, groups, eps, layout, st);
  }
  if (vec)
    groupnorm_stats_kernel<T, true><<<B * F * groups, 256, 0, st>>>((const T*)x, stats, C, F, HW, groups, eps);
  else
    groupnorm_stats_kernel<T, false><<<B * F * groups, 256, 0, st>>>((const T*)x, stats, C, F, HW, groups, eps);
  CVD_LAUNCH_CHECK("groupnorm_stats_kernel");
  dim3 grid((HW + kTileP - 1) / kTileP, (C + kTileC - 1) / kTileC, B * F);
  video_to_tokens_kernel<T, 0><<<grid, 256, 0, st>>>((const T*)x, stats, gamma, beta, nullptr, (__nv_bfloat16*)y, C, F,
                                                     HW, groups, layout);
  CVD_LAUNCH_CHECK("video_to_tokens_kernel");
  return CVD_OK;
}

}  // namespace cvd

using namespace cvd;

extern "C" int cvd_groupnorm_tokens_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                                        float* stats, int B, int C, int F, int HW, int groups, float eps, int layout,
                                        cvd_stream_t stream) {
  CVD_REQUIRE(x && gamma && beta && y && stats, "cvd_groupnorm_tokens_fwd: null pointer");
  CVD_REQUIRE(B > 0 && C > 0 && F > 0 && HW > 0 && groups > 0 && C % groups == 0 && C % 2 == 0,
              "cvd_groupnorm_tokens_fwd: bad shape B=%d C=%d F=%d HW=%d groups=%d", B, C, F, HW, groups);
  CVD_REQUIRE((long long)B * F <= 65535, "cvd_groupnorm_tokens_fwd: B*F=%lld exceeds the grid limit", (long long)B * F);
  CVD_REQUIRE(layout == CVD_LAYOUT_BF_HW || layout == CVD_LAYOUT_BHW_F, "cvd_groupnorm_tokens_fwd: layout %d", layout);
  cudaStream_t st = (cudaStream_t)stream;
  switch (x_dtype) {
    case CVD_DTYPE_F32: return launch_groupnorm_tokens<float>(x, gamma, beta, y, stats, B, C, F, HW, groups, eps, layout, st);
    case CVD_DTYPE_BF16: return launch_groupnorm_tokens<__nv_bfloat16>(x, gamma, beta, y, stats, B, C, F, HW, groups, eps, layout, st);
    case CVD_DTYPE_F16: return launch_groupnorm_tokens<__half>(x, gamma, beta, y, stats, B, C, F, HW, groups, eps, layout, st);
  }
  return fail(CVD_E_INVALID, "cvd_groupnorm_tokens_fwd: x_dtype %d not supported", x_dtype);
}

extern "C" int cvd_add_pose_tokens_fwd(const void* x_tok, const void* pose, int pose_dtype, void* out_tok, int B, int C,
                                       int F, int HW, cvd_stream_t stream) {
  CVD_REQUIRE(x_tok && pose && out_tok, "cvd_add_pose_tokens_fwd: null pointer");
  CVD_REQUIRE(B > 0 && C > 0 && C % 2 == 0 && F > 0 && HW > 0 && (long long)B * F <= 65535,
              "cvd_add_pose_tokens_fwd: bad shape B=%d C=%d F=%d HW=%d", B, C, F, HW);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((HW + kTileP - 1) / kTileP, (C + kTileC - 1) / kTileC, B * F);
  const __nv_bfloat16* xin = (const __nv_bfloat16*)x_tok;
  __nv_bfloat16* o = (__nv_bfloat16*)out_tok;
  if (pose_dtype == CVD_DTYPE_F32)
    video_to_tokens_kernel<float, 1><<<grid, 256, 0, st>>>((const float*)pose, nullptr, nullptr, nullptr, xin, o, C, F, HW, 1, CVD_LAYOUT_BHW_F);
  else if (pose_dtype == CVD_DTYPE_BF16)
    video_to_tokens_kernel<__nv_bfloat16, 1><<<grid, 256, 0, st>>>((const __nv_bfloat16*)pose, nullptr, nullptr, nullptr, xin, o, C, F, HW, 1, CVD_LAYOUT_BHW_F);
  else if (pose_dtype == CVD_DTYPE_F16)
    video_to_tokens_kernel<__half, 1><<<grid, 256, 0, st>>>((const __half*)pose, nullptr, nullptr, nullptr, xin, o, C, F, HW, 1, CVD_LAYOUT_BHW_F);
  else
    return fail(CVD_E_INVALID, "cvd_add_pose_tokens_fwd: pose_dtype %d not supported", pose_dtype);
  CVD_LAUNCH_CHECK("video_to_tokens_kernel<add>");
  return CVD_OK;
}

extern "C" int cvd_tokens_to_video_fwd(const void* tok, const float* bias, const void* residual, void* out, int io_dtype,
                                       int B, int C, int F, int HW, int layout, cvd_stream_t stream) {
  CVD_REQUIRE(tok && residual && out, "cvd_tokens_to_video_fwd: null pointer");
  CVD_REQUIRE(B > 0 && C > 0 && C % 2 == 0 && F > 0 && HW > 0 && (long long)B * F <= 65535,
              "cvd_tokens_to_video_fwd: bad shape B=%d C=%d F=%d HW=%d", B, C, F, HW);
  CVD_REQUIRE(layout == CVD_LAYOUT_BF_HW || layout == CVD_LAYOUT_BHW_F, "cvd_tokens_to_video_fwd: layout %d", layout);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((HW + kTileP - 1) / kTileP, (C + kTileC - 1) / kTileC, B * F);
  const __nv_bfloat16* t = (const __nv_bfloat16*)tok;
  // full-row version where it has enough CTAs to fill the machine (measured: 55 vs 66 us at 32x32, but 46 vs 36 us at
  // 16x16 and 74 vs 19 us at 8x8, where B*F*HW/32 CTAs are too few)
  if (io_dtype == CVD_DTYPE_F32 && C % 64 == 0 && HW % kTileP == 0 && (long long)B * F * (HW / kTileP) >= 2048 &&
      (((uintptr_t)tok) & 15) == 0) {
    const size_t smem = (size_t)kTileP * (C / 2 + 1) * sizeof(uint32_t);
    CVD_CUDA(ensure_dynamic_smem((const void*)tokens_to_video_rows_kernel, smem));
    tokens_to_video_rows_kernel<<<(unsigned)(B * F * (HW / kTileP)), 256, smem, st>>>(t, bias, (const float*)residual, (float*)out,
                                                                                  C, F, HW, layout);
    CVD_LAUNCH_CHECK("tokens_to_video_rows_kernel");
    return CVD_OK;
  }
  if (io_dtype == CVD_DTYPE_F32)
    tokens_to_video_kernel<float><<<grid, 256, 0, st>>>(t, bias, (const float*)residual, (float*)out, C, F, HW, layout);
  else if (io_dtype == CVD_DTYPE_BF16)
    tokens_to_video_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(t, bias, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, C, F, HW, layout);
  else if (io_dtype == CVD_DTYPE_F16)
    tokens_to_video_kernel<__half><<<grid, 256, 0, st>>>(t, bias, (const __half*)residual, (__half*)out, C, F, HW, layout);
  else
    return fail(CVD_E_INVALID, "cvd_tokens_to_video_fwd: io_dtype %d not supported", io_dtype);
  CVD_LAUNCH_CHECK("tokens_to_video_kernel");
  return CVD_OK;
}

extern "C" int cvd_layernorm_fwd(const float* x, const float* gamma, const float* beta, const float* pe, void* y,
                                 int tokens, int C, int seq_len, float eps, cvd_stream_t stream) {
  CVD_REQUIRE(x && gamma && beta && y, "cvd_layernorm_fwd: null pointer");
  CVD_REQUIRE(tokens > 0, "cvd_layernorm_fwd: tokens=%d", tokens);
  CVD_REQUIRE(!pe || seq_len > 0, "cvd_layernorm_fwd: seq_len must be positive with a positional table");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((tokens + 7) / 8);
  __nv_bfloat16* o = (__nv_bfloat16*)y;
  switch (C) {
    case 64: layernorm_kernel<1><<<grid, 256, 0, st>>>(x, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 320: layernorm_kernel<5><<<grid, 256, 0, st>>>(x, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 640: layernorm_kernel<10><<<grid, 256, 0, st>>>(x, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 1280: layernorm_kernel<20><<<grid, 256, 0, st>>>(x, gamma, beta, pe, o, tokens, seq_len, eps); break;
    default: return fail(CVD_E_INVALID, "cvd_layernorm_fwd: C=%d not supported (64, 320, 640, 1280)", C);
  }
  CVD_LAUNCH_CHECK("layernorm_kernel");
  return CVD_OK;
}

extern "C" int cvd_residual_layernorm_fwd(float* x, const void* y, const float* bias, float scale, const float* gamma,
                                          const float* beta, const float* pe, void* n_out, int tokens, int C, int seq_len,
                                          float eps, cvd_stream_t stream) {
  CVD_REQUIRE(x && gamma && beta && n_out, "cvd_residual_layernorm_fwd: null pointer");
  CVD_REQUIRE(tokens > 0 && seq_len > 0, "cvd_residual_layernorm_fwd: tokens=%d seq_len=%d", tokens, seq_len);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((tokens + 7) / 8);
  const __nv_bfloat16* yy = (const __nv_bfloat16*)y;
  __nv_bfloat16* o = (__nv_bfloat16*)n_out;
  switch (C) {
    case 64: residual_layernorm_kernel<1><<<grid, 256, 0, st>>>(x, yy, bias, scale, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 320: residual_layernorm_kernel<5><<<grid, 256, 0, st>>>(x, yy, bias, scale, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 640: residual_layernorm_kernel<10><<<grid, 256, 0, st>>>(x, yy, bias, scale, gamma, beta, pe, o, tokens, seq_len, eps); break;
    case 1280: residual_layernorm_kernel<20><<<grid, 256, 0, st>>>(x, yy, bias, scale, gamma, beta, pe, o, tokens, seq_len, eps); break;
    default: return fail(CVD_E_INVALID, "cvd_residual_layernorm_fwd: C=%d not supported (64, 320, 640, 1280)", C);
  }
  CVD_LAUNCH_CHECK("residual_layernorm_kernel");
  return CVD_OK;
}

extern "C" int cvd_geglu_fwd(const void* hg, void* out, int tokens, int inner, cvd_stream_t stream) {
  CVD_REQUIRE(hg && out, "cvd_geglu_fwd: null pointer");
  CVD_REQUIRE(tokens > 0 && inner > 0 && inner % 8 == 0, "cvd_geglu_fwd: tokens=%d inner=%d (inner must be a multiple of 8)", tokens, inner);
  const size_t n = (size_t)tokens * (inner / 8);
  geglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)hg, (uint4*)out, (size_t)tokens, inner / 8);
  CVD_LAUNCH_CHECK("geglu_kernel");
  return CVD_OK;
}

extern "C" int cvd_residual_add_fwd(float* x, const void* y, const float* bias, float scale, void* x_bf16, int tokens,
                                    int C, cvd_stream_t stream) {
  CVD_REQUIRE(x && y, "cvd_residual_add_fwd: null pointer");
  CVD_REQUIRE(tokens > 0 && C > 0 && C % 4 == 0, "cvd_residual_add_fwd: tokens=%d C=%d (C must be a multiple of 4)", tokens, C);
  const size_t n4 = (size_t)tokens * (C / 4);
  residual_add_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (float4*)x, (const uint2*)y, (const float4*)bias, scale, (uint2*)x_bf16, n4, C / 4);
  CVD_LAUNCH_CHECK("residual_add_kernel");
  return CVD_OK;
}

extern "C" int cvd_scale_add_fwd(const void* a, const void* b, float scale, void* out, long long n, cvd_stream_t stream) {
  CVD_REQUIRE(a && b && out, "cvd_scale_add_fwd: null pointer");
  CVD_REQUIRE(n > 0 && n % 8 == 0, "cvd_scale_add_fwd: n=%lld must be a positive multiple of 8", n);
  const size_t n8 = (size_t)n / 8;
  scale_add_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)b, scale, (uint4*)out, n8);
  CVD_LAUNCH_CHECK("scale_add_kernel");
  return CVD_OK;
}
