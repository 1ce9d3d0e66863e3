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

// ---------------------------------------------------------------------------------------
// F == 16 (the reference's clip length): warp-level tensor-core version.  One CTA per pixel, one
// warp per head.  q/k/v slabs [16, C] go global -> shared with 16-byte cp.async (fully coalesced,
// no register staging); S = Q K^T (16x16xd) and O = P V (16xdx16) are mma.sync m16n8k16 / k8 on
// ldmatrix fragments, the softmax runs on the accumulator fragments (quad shuffles), P never
// leaves registers, and O returns through the q slab for a coalesced 16-byte store.
// The math is ~1 KFLOP per output element short of nothing: the kernel is HBM-bound
// (4 * T * 16 * C * 2 bytes), which is the only reason to bother with the tensor cores here --
// the scalar version above spends 5-6x the HBM time in shared-memory-bound inner loops.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t pack_bf162(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <int D>
__global__ void __launch_bounds__(256)
temporal_attn16_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                       const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int heads) {
  constexpr int F = 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = heads * D;
  const int pitch = C + 8;                        // rows 16-byte aligned, consecutive rows 4 banks apart (ldmatrix conflict-free)
  const uint32_t sq = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t sk = sq + F * pitch * 2, sv = sk + F * pitch * 2;
  const size_t base = (size_t)blockIdx.x * F * C;
  {
    const int c8 = C / 8;
    for (int i = threadIdx.x; i < F * c8; i += blockDim.x) {
      const uint32_t o = ((i / c8) * pitch + (i % c8) * 8) * 2;
      const size_t g = base + (size_t)i * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sq + o), "l"(q + g));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sk + o), "l"(k + g));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sv + o), "l"(v + g));
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = rsqrtf((float)D) * 1.4426950408889634f;
  for (int h = warp; h < heads; h += nwarps) {
    const uint32_t hq = sq + h * D * 2, hk = sk + h * D * 2, hv = sv + h * D * 2;
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};   // key frames 0-7 | 8-15
#pragma unroll
    for (int k0 = 0; k0 + 16 <= D; k0 += 16) {
      uint32_t a[4], b[4];
      ldsm_x4(hq + (((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + k0 + (lane >> 4) * 8) * 2, a);
      ldsm_x4(hk + (((lane & 7) + (lane >> 4) * 8) * pitch + k0 + ((lane >> 3) & 1) * 8) * 2, b);
      mma_16816(s0, a, b[0], b[1]);
      mma_16816(s1, a, b[2], b[3]);
    }
    if (D % 16 == 8) {
      constexpr int k0 = D - 8;
      uint32_t a[2], b[2];
      ldsm_x2(hq + ((lane & 15) * pitch + k0) * 2, a);
      ldsm_x2(hk + ((lane & 15) * pitch + k0) * 2, b);
      mma_1688(s0, a[0], a[1], b[0]);
      mma_1688(s1, a[0], a[1], b[1]);
    }
    // softmax over the 16 key frames: rows g (c0,c1) and g+8 (c2,c3); a row is spread over the 4 lanes of a quad
    float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * scale_log2, o1 = m1 * scale_log2;
    s0[0] = exp2f(fmaf(s0[0], scale_log2, -o0)); s0[1] = exp2f(fmaf(s0[1], scale_log2, -o0));
    s1[0] = exp2f(fmaf(s1[0], scale_log2, -o0)); s1[1] = exp2f(fmaf(s1[1], scale_log2, -o0));
    s0[2] = exp2f(fmaf(s0[2], scale_log2, -o1)); s0[3] = exp2f(fmaf(s0[3], scale_log2, -o1));
    s1[2] = exp2f(fmaf(s1[2], scale_log2, -o1)); s1[3] = exp2f(fmaf(s1[3], scale_log2, -o1));
    float l0 = s0[0] + s0[1] + s1[0] + s1[1], l1 = s0[2] + s0[3] + s1[2] + s1[3];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // P as the A operand (16 x 16): the accumulator layout of S is the A layout.  P is split into a bf16 head and a
    // bf16 remainder (two MMAs per tile, free in an HBM-bound kernel) so that the probabilities keep ~16 mantissa
    // bits, as in the fp32 softmax of the reference.
    uint32_t p[4], pl[4];
    {
      const float pf[8] = {s0[0] * i0, s0[1] * i0, s0[2] * i1, s0[3] * i1, s1[0] * i0, s1[1] * i0, s1[2] * i1, s1[3] * i1};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hi = __floats2bfloat162_rn(pf[2 * i], pf[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hi);
        p[i] = *reinterpret_cast<const uint32_t*>(&hi);
        pl[i] = pack_bf162(pf[2 * i] - hf.x, pf[2 * i + 1] - hf.y);
      }
    }
    __syncwarp();        // every lane has read its Q fragments: the head's q columns may be overwritten with O
#pragma unroll
    for (int c0 = 0; c0 + 16 <= D; c0 += 16) {
      uint32_t b[4];
      ldsm_x4_t(hv + (((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + c0 + (lane >> 4) * 8) * 2, b);
      float oa[4] = {0.f, 0.f, 0.f, 0.f}, ob[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(oa, p, b[0], b[1]);
      mma_16816(ob, p, b[2], b[3]);
      mma_16816(oa, pl, b[0], b[1]);
      mma_16816(ob, pl, b[2], b[3]);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(smem_raw) + h * D + c0 + 2 * t;
      *reinterpret_cast<uint32_t*>(dst + g * pitch) = pack_bf162(oa[0], oa[1]);
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * pitch) = pack_bf162(oa[2], oa[3]);
      *reinterpret_cast<uint32_t*>(dst + g * pitch + 8) = pack_bf162(ob[0], ob[1]);
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * pitch + 8) = pack_bf162(ob[2], ob[3]);
    }
    if (D % 16 == 8) {
      constexpr int c0 = D - 8;
      uint32_t b[2];
      ldsm_x2_t(hv + ((lane & 15) * pitch + c0) * 2, b);
      float oa[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(oa, p, b[0], b[1]);
      mma_16816(oa, pl, b[0], b[1]);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(smem_raw) + h * D + c0 + 2 * t;
      *reinterpret_cast<uint32_t*>(dst + g * pitch) = pack_bf162(oa[0], oa[1]);
      *reinterpret_cast<uint32_t*>(dst + (g + 8) * pitch) = pack_bf162(oa[2], oa[3]);
    }
  }
  __syncthreads();
  {
    uint4* go = reinterpret_cast<uint4*>(out + base);
    const int c8 = C / 8;
    for (int i = threadIdx.x; i < F * c8; i += blockDim.x)
      go[i] = *reinterpret_cast<const uint4*>(smem_raw + ((i / c8) * pitch + (i % c8) * 8) * 2);
  }
}

template <int D>
static int launch16(const void* q, const void* k, const void* v, void* out, int T, int heads, cudaStream_t st) {
  const int C = heads * D;
  const int warps = heads < 8 ? heads : 8;
  const size_t smem = (size_t)3 * 16 * (C + 8) * sizeof(__nv_bfloat16);
  CVD_REQUIRE(smem <= 227 * 1024, "cvd_temporal_attn_fwd: C=%d needs %zu bytes of shared memory", C, smem);
  CVD_CUDA(ensure_dynamic_smem((const void*)temporal_attn16_kernel<D>, smem));
  temporal_attn16_kernel<D><<<T, warps * 32, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                         (const __nv_bfloat16*)v, (__nv_bfloat16*)out, heads);
  CVD_LAUNCH_CHECK("temporal_attn16_kernel");
  return CVD_OK;
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
  if (F == 16) {   // the reference's clip length: tensor-core kernel for the head sizes of the SD-1.5 UNet
    if (head_dim == 40) return launch16<40>(q, k, v, out, T, heads, (cudaStream_t)stream);
    if (head_dim == 80) return launch16<80>(q, k, v, out, T, heads, (cudaStream_t)stream);
    if (head_dim == 160) return launch16<160>(q, k, v, out, T, heads, (cudaStream_t)stream);
  }
  const int warps = heads < 8 ? heads : 8;
  const size_t smem = (size_t)3 * F * (C + 2) * sizeof(__nv_bfloat16) + (size_t)warps * F * (F + 1) * sizeof(float);
  CVD_REQUIRE(smem <= 227 * 1024, "cvd_temporal_attn_fwd: F=%d C=%d needs %zu bytes of shared memory", F, C, smem);
  CVD_CUDA(ensure_dynamic_smem((const void*)temporal_attn_kernel, smem));
  temporal_attn_kernel<<<T, warps * 32, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, F, heads, head_dim);
  CVD_LAUNCH_CHECK("temporal_attn_kernel");
  return CVD_OK;
}
