// Temporal self-attention of the AnimateDiff motion module: every pixel attends over its
// F (<= 32, reference: 16) frames (animatediff/models/motion_module.py:509-549 with the
// processors at attention_processor.py:84-94 and :346-357; both compute
// softmax(q k^T / sqrt(d)) v per head).  The contraction is 16x16xd per (pixel, head) --
// far too small for tensor cores -- so the kernel is HBM-bound: one CTA per pixel loads the
// pixel's q/k/v rows [F, C] with 16-byte coalesced loads into shared memory, one warp per
// head does the math in fp32, and the result leaves through the same coalesced path.
#include "common.cuh"

namespace cvd {

__global__ void __launch_bounds__(256)
temporal_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int F, int heads, int d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = heads * d;
  const int pitch = C + 2;                        // odd number of 32-bit words: conflict-free row stride
  const int rowlen = F * C;                       // elements of one pixel's [F, C] slab in HBM
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sk = sq + F * pitch;
  __nv_bfloat16* sv = sk + F * pitch;
  float* sp = reinterpret_cast<float*>(sv + F * pitch);   // [warps][F][F+1] probabilities
  const size_t base = (size_t)blockIdx.x * rowlen;
  {
    const uint4* gq = reinterpret_cast<const uint4*>(q + base);
    const uint4* gk = reinterpret_cast<const uint4*>(k + base);
    const uint4* gv = reinterpret_cast<const uint4*>(v + base);
    const int c8 = C / 8;
    for (int i = threadIdx.x; i < rowlen / 8; i += blockDim.x) {
      const int o = (i / c8) * pitch + (i % c8) * 8;
      const uint4 a = gq[i], b = gk[i], c = gv[i];
      uint32_t* dq = reinterpret_cast<uint32_t*>(sq + o);
      uint32_t* dk = reinterpret_cast<uint32_t*>(sk + o);
      uint32_t* dv = reinterpret_cast<uint32_t*>(sv + o);
      dq[0] = a.x; dq[1] = a.y; dq[2] = a.z; dq[3] = a.w;
      dk[0] = b.x; dk[1] = b.y; dk[2] = b.z; dk[3] = b.w;
      dv[0] = c.x; dv[1] = c.y; dv[2] = c.z; dv[3] = c.w;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const float scale = rsqrtf((float)d);
  float* P = sp + warp * F * (F + 1);
  for (int h = warp; h < heads; h += nwarps) {
    // scores: F*F pairs spread over the 32 lanes
    for (int e = lane; e < F * F; e += 32) {
      const int i = e / F, j = e % F;
      const __nv_bfloat162* qi = reinterpret_cast<const __nv_bfloat162*>(sq + i * pitch + h * d);
      const __nv_bfloat162* kj = reinterpret_cast<const __nv_bfloat162*>(sk + j * pitch + h * d);
      float acc = 0.f;
      for (int c = 0; c < d / 2; ++c) {
        const float2 a = __bfloat1622float2(qi[c]), b = __bfloat1622float2(kj[c]);
        acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
      }
      P[i * (F + 1) + j] = acc * scale;
    }
    __syncwarp();
    if (lane < F) {   // one lane per query row: softmax over its F scores
      float* row = P + lane * (F + 1);
      float m = -INFINITY;
      for (int j = 0; j < F; ++j) m = fmaxf(m, row[j]);
      float s = 0.f;
      for (int j = 0; j < F; ++j) {
        const float e = __expf(row[j] - m);
        row[j] = e;
        s += e;
      }
      const float inv = 1.f / s;
      for (int j = 0; j < F; ++j) row[j] *= inv;
    }
    __syncwarp();
    // out[i, c] = sum_j P[i,j] v[j,c]; the q slab of this head is dead now: reuse it as output
    for (int e = lane; e < F * (d / 2); e += 32) {
      const int i = e / (d / 2), c = e % (d / 2);
      float2 acc = make_float2(0.f, 0.f);
      for (int j = 0; j < F; ++j) {
        const float p = P[i * (F + 1) + j];
        const float2 vv = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(sv + j * pitch + h * d)[c]);
        acc.x = fmaf(p, vv.x, acc.x);
        acc.y = fmaf(p, vv.y, acc.y);
      }
      reinterpret_cast<__nv_bfloat162*>(sq + i * pitch + h * d)[c] = __floats2bfloat162_rn(acc.x, acc.y);
    }
    __syncwarp();
  }
  __syncthreads();
  uint4* go = reinterpret_cast<uint4*>(out + base);
  const int c8 = C / 8;
  for (int i = threadIdx.x; i < rowlen / 8; i += blockDim.x) {
    const uint32_t* so = reinterpret_cast<const uint32_t*>(sq + (i / c8) * pitch + (i % c8) * 8);
    go[i] = make_uint4(so[0], so[1], so[2], so[3]);
  }
}

}  // namespace cvd

extern "C" int cvd_temporal_attn_fwd(const void* q, const void* k, const void* v, void* out, int T, int F, int heads,
                                     int head_dim, cvd_stream_t stream) {
  using namespace cvd;
  CVD_REQUIRE(q && k && v && out, "cvd_temporal_attn_fwd: null pointer");
  CVD_REQUIRE(T > 0 && F > 0 && F <= 32 && heads > 0 && head_dim > 0 && head_dim % 2 == 0,
              "cvd_temporal_attn_fwd: T=%d F=%d heads=%d head_dim=%d", T, F, heads, head_dim);
  const int C = heads * head_dim;
  CVD_REQUIRE(C % 8 == 0, "cvd_temporal_attn_fwd: heads*head_dim must be a multiple of 8");
  const int warps = heads < 8 ? heads : 8;
  const size_t smem = (size_t)3 * F * (C + 2) * sizeof(__nv_bfloat16) + (size_t)warps * F * (F + 1) * sizeof(float);
  CVD_REQUIRE(smem <= 227 * 1024, "cvd_temporal_attn_fwd: F=%d C=%d needs %zu bytes of shared memory", F, C, smem);
  static size_t configured_smem = 0;   // grow-only; one process per GPU
  if (smem > configured_smem) {
    CVD_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_smem = smem;
  }
  temporal_attn_kernel<<<T, warps * 32, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, F, heads, head_dim);
  CVD_LAUNCH_CHECK("temporal_attn_kernel");
  return CVD_OK;
}
