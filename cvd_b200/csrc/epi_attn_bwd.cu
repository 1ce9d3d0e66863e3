// Backward of the epipolar cross-view attention core (train_epi_control.py runs
// loss.backward() through EpiAttnProcessor, attention_processor.py:573-619; the bias has no gradient,
// epi_module.py:329,393):
//
//     P = softmax_k(q k^T / sqrt(d) + bias)        (recomputed from the saved base-2 log-sum-exp, never stored)
//     dV = P^T dO,   dP = dO V^T,   dS = P o (dP - rowsum(dO o O)),   dQ = dS K / sqrt(d),   dK = dS^T Q / sqrt(d)
//
// First CUDA path for this row: correctness first, warp-level tensor cores (mma.sync m16n8k16, bf16 operands, fp32
// accumulation), no tcgen05 yet.  One CTA owns a block of BLK keys of one (row, head) and a 40-wide slice of the head
// dimension of the three outputs; it walks the query blocks, recomputes S and dP over the full head dimension,
// keeps its dK / dV slices in registers and adds dQ slices to global memory.  The epipolar bias is re-evaluated in
// registers from the 12-float row table exactly as in the forward kernels.  Outputs are fp32 and ACCUMULATED with
// atomics (the caller zero-fills): K/V rows are addressed through the partner map, which need not be injective.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cvd {
namespace bwd {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// fp32 pair reduction into global memory (REDG.ADD.F32x2): half the reduction requests of two scalar atomics
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// delta[r, h, n] = sum_c dO[r, n, h*d + c] * O[r, n, h*d + c]      (one warp per (r, n, h))
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                      long long items, int N, int H, int D) {
  const long long it = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (it >= items) return;
  const int lane = threadIdx.x & 31;
  const int h = (int)(it % H);
  const long long rn = it / H;
  const size_t base = (size_t)rn * H * D + (size_t)h * D;
  float acc = 0.f;
  for (int c = lane * 2; c < D; c += 64) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + base + c));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(d_o + base + c));
    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[((size_t)(rn / N) * H + h) * N + (size_t)(rn % N)] = acc;
}

template <int D, int BLK>
struct CfgB {
  static constexpr int kD16 = (D + 15) / 16 * 16;
  static constexpr int kDP = kD16 + 8;              // pitch of the row-major tiles (elements): 16-byte rows, 4-bank shift
  static constexpr int kTP = BLK + 8;               // pitch of the transposed tiles
  static constexpr int kWarps = BLK / 16;
  // head-dim slice of dQ / dK / dV per CTA: S and dP are recomputed once per slice, so wide heads take 80-wide slices
  // (d = 160: 2 instead of 4 recomputations, d = 80: none); the accumulators of an 80-wide slice are 120 registers
  static constexpr int kDH = D % 80 == 0 ? 80 : 40;
  static constexpr int kChunks = D / kDH;
  static_assert(D % kDH == 0, "head_dim must be a multiple of 40");
  static constexpr size_t kSmem = (size_t)(4 * BLK * kDP + 3 * kDH * kTP + 2 * BLK * kTP) * sizeof(__nv_bfloat16);
};

// rows [row0, row0 + BLK) of a [rows, C] bf16 matrix, columns [h*D, h*D + D) -> dst [BLK][DP] (pad columns zero) and,
// if dst_t != nullptr, columns [c0, c0 + 40) transposed -> dst_t [40][TP]
template <int D, int BLK>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ src, size_t row0, int C, int h, int c0,
                                          __nv_bfloat16* dst, __nv_bfloat16* dst_t) {
  using Cf = CfgB<D, BLK>;
  constexpr int kVec = Cf::kD16 / 8;
  constexpr int kDH = Cf::kDH;
  for (int i = threadIdx.x; i < BLK * kVec; i += Cf::kWarps * 32) {
    const int row = i / kVec, v = i % kVec;
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (v * 8 < D) t = *reinterpret_cast<const uint4*>(src + (row0 + row) * (size_t)C + (size_t)h * D + v * 8);
    *reinterpret_cast<uint4*>(dst + row * Cf::kDP + v * 8) = t;
    if (dst_t != nullptr && v * 8 >= c0 && v * 8 < c0 + kDH) {
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&t);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst_t[(v * 8 - c0 + j) * Cf::kTP + row] = e[j];
    }
  }
}

template <int D, int BLK>
__global__ void __launch_bounds__(CfgB<D, BLK>::kWarps * 32)
epi_attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                    const __nv_bfloat16* __restrict__ v, const float* __restrict__ rowp, const int* __restrict__ partner,
                    const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse, const float* __restrict__ delta,
                    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int N, int H, int feat, float step) {
  using Cf = CfgB<D, BLK>;
  constexpr int kDH = Cf::kDH;
  constexpr int DP = Cf::kDP, TP = Cf::kTP, NT = BLK / 8, KS = Cf::kD16 / 16, OT = kDH / 8;
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(bwd_smem);
  __nv_bfloat16* sdO = sQ + BLK * DP;
  __nv_bfloat16* sK = sdO + BLK * DP;
  __nv_bfloat16* sV = sK + BLK * DP;
  __nv_bfloat16* sQt = sV + BLK * DP;          // [40][TP]  Q^T slice
  __nv_bfloat16* sdOt = sQt + kDH * TP;        // [40][TP]  dO^T slice
  __nv_bfloat16* sKt = sdOt + kDH * TP;        // [40][TP]  K^T slice
  __nv_bfloat16* sPt = sKt + kDH * TP;         // [BLK keys][TP]  P^T
  __nv_bfloat16* sdSt = sPt + BLK * TP;        // [BLK keys][TP]  dS^T

  const int C = H * D;
  const int jb = blockIdx.x / Cf::kChunks, c0 = (blockIdx.x % Cf::kChunks) * kDH;   // key block, head-dim slice
  const int h = blockIdx.y, r = blockIdx.z;
  const int rr = partner[r];                     // K / V row of this query row (single partner)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf((float)D), scale2 = scale * kLog2e;
  const float off = (step - 1.0f) * 0.5f;

  load_tile<D, BLK>(k, (size_t)rr * N + (size_t)jb * BLK, C, h, c0, sK, sKt);
  load_tile<D, BLK>(v, (size_t)rr * N + (size_t)jb * BLK, C, h, c0, sV, nullptr);

  // row table of this (r): effective line matrix, band, weight
  float Pm[11];
  if (rowp != nullptr) {
#pragma unroll
    for (int i = 0; i < 11; ++i) Pm[i] = rowp[(size_t)r * CVD_ROWP_STRIDE + i];
  }
  const float bandp = rowp != nullptr ? Pm[9] * Pm[10] * kLog2e : 0.f;

  float dk_acc[OT][4], dv_acc[OT][4];
#pragma unroll
  for (int i = 0; i < OT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk_acc[i][e] = dv_acc[i][e] = 0.f;

  const int nqb = N / BLK;
  for (int ib = 0; ib < nqb; ++ib) {
    __syncthreads();                             // previous iteration's readers of sQ / sdO / sPt / sdSt are done
    load_tile<D, BLK>(q, (size_t)r * N + (size_t)ib * BLK, C, h, c0, sQ, sQt);
    load_tile<D, BLK>(d_o, (size_t)r * N + (size_t)ib * BLK, C, h, c0, sdO, sdOt);
    __syncthreads();

    // ---- S = Q K^T and dP = dO V^T for this warp's 16 query rows x BLK keys
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
    const __nv_bfloat16* qa = sQ + (warp * 16 + g) * DP + 2 * t;
    const __nv_bfloat16* da = sdO + (warp * 16 + g) * DP + 2 * t;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t a0 = lds32(qa + ks * 16), a1 = lds32(qa + 8 * DP + ks * 16), a2 = lds32(qa + ks * 16 + 8),
                     a3 = lds32(qa + 8 * DP + ks * 16 + 8);
      const uint32_t e0 = lds32(da + ks * 16), e1 = lds32(da + 8 * DP + ks * 16), e2 = lds32(da + ks * 16 + 8),
                     e3 = lds32(da + 8 * DP + ks * 16 + 8);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const __nv_bfloat16* kb = sK + (nt * 8 + g) * DP + ks * 16 + 2 * t;
        const __nv_bfloat16* vb = sV + (nt * 8 + g) * DP + ks * 16 + 2 * t;
        mma16816(s[nt], a0, a1, a2, a3, lds32(kb), lds32(kb + 8));
        mma16816(dp[nt], e0, e1, e2, e3, lds32(vb), lds32(vb + 8));
      }
    }

    // ---- P = exp2(S * scale2 + bias2 - lse2), dS = P o (dP - delta); rows g and g + 8 of this warp's strip
    const int qrow[2] = {ib * BLK + warp * 16 + g, ib * BLK + warp * 16 + g + 8};
    float As[2] = {0.f, 0.f}, Bs[2] = {0.f, 0.f}, base[2] = {0.f, 0.f}, l2[2], dl[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      l2[i] = lse[((size_t)r * H + h) * N + qrow[i]];
      dl[i] = delta[((size_t)r * H + h) * N + qrow[i]];
      if (rowp != nullptr) {
        const float xq = pixel_centre(qrow[i] % feat, step), yq = pixel_centre(qrow[i] / feat, step);
        const float a = fmaf(Pm[1], yq, Pm[0] * xq) + Pm[2];
        const float b = fmaf(Pm[4], yq, Pm[3] * xq) + Pm[5];
        const float c = fmaf(Pm[7], yq, Pm[6] * xq) + Pm[8];
        const float k1 = Pm[10] * kLog2e / (sqrtf(a * a + b * b) + 1e-6f);
        As[i] = a * k1 * step;
        Bs[i] = b * k1 * step;
        base[i] = fmaf(a * k1, off, fmaf(b * k1, off, c * k1));
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kl = nt * 8 + 2 * t + (e & 1);           // key inside the block
        const int key = jb * BLK + kl;
        const int i = e >> 1;
        float x = fmaf(s[nt][e], scale2, -l2[i]);
        if (rowp != nullptr) {
          const float tt = fmaf(As[i], (float)(key % feat), fmaf(Bs[i], (float)(key / feat), base[i]));
          float bias = fminf(bandp - fabsf(tt), 0.f);
          if (!(fabsf(bias) <= 3.0e38f)) bias = 0.f;       // NaN / inf geometry -> 0 (epi_module.py:404-407)
          x += bias;
        }
        const float p = ptx::ex2(x);
        const float ds = p * (dp[nt][e] - dl[i]);
        s[nt][e] = ds;                                     // s now holds dS (A operand of dQ)
        const int ql = warp * 16 + g + 8 * i;              // query inside the block
        sPt[kl * TP + ql] = __float2bfloat16_rn(p);
        sdSt[kl * TP + ql] = __float2bfloat16_rn(ds);
      }
    }

    // ---- dQ[16 x 40] = dS[16 x BLK] K[BLK x 40-slice] / sqrt(d)  -> global (atomic)
    {
      float dqa[OT][4];
#pragma unroll
      for (int i = 0; i < OT; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) dqa[i][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < BLK / 16; ++kk) {
        const uint32_t a0 = packbf(s[2 * kk][0], s[2 * kk][1]), a1 = packbf(s[2 * kk][2], s[2 * kk][3]),
                       a2 = packbf(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = packbf(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int ot = 0; ot < OT; ++ot) {
          const __nv_bfloat16* b = sKt + (ot * 8 + g) * TP + kk * 16 + 2 * t;
          mma16816(dqa[ot], a0, a1, a2, a3, lds32(b), lds32(b + 8));
        }
      }
#pragma unroll
      for (int ot = 0; ot < OT; ++ot) {
        float* d0 = dq + ((size_t)r * N + qrow[0]) * C + (size_t)h * D + c0 + ot * 8 + 2 * t;
        float* d1 = dq + ((size_t)r * N + qrow[1]) * C + (size_t)h * D + c0 + ot * 8 + 2 * t;
        red_add2(d0, dqa[ot][0] * scale, dqa[ot][1] * scale);
        red_add2(d1, dqa[ot][2] * scale, dqa[ot][3] * scale);
      }
    }
    __syncthreads();                             // P^T and dS^T of all warps are in shared memory

    // ---- dV[16 keys x 40] += P^T[16 x BLK q] dO[BLK q x 40-slice];  dK likewise with dS^T and Q
    const __nv_bfloat16* pa = sPt + (warp * 16 + g) * TP + 2 * t;
    const __nv_bfloat16* sa = sdSt + (warp * 16 + g) * TP + 2 * t;
#pragma unroll
    for (int kk = 0; kk < BLK / 16; ++kk) {
      const uint32_t a0 = lds32(pa + kk * 16), a1 = lds32(pa + 8 * TP + kk * 16), a2 = lds32(pa + kk * 16 + 8),
                     a3 = lds32(pa + 8 * TP + kk * 16 + 8);
      const uint32_t e0 = lds32(sa + kk * 16), e1 = lds32(sa + 8 * TP + kk * 16), e2 = lds32(sa + kk * 16 + 8),
                     e3 = lds32(sa + 8 * TP + kk * 16 + 8);
#pragma unroll
      for (int ot = 0; ot < OT; ++ot) {
        const __nv_bfloat16* bo = sdOt + (ot * 8 + g) * TP + kk * 16 + 2 * t;
        const __nv_bfloat16* bq = sQt + (ot * 8 + g) * TP + kk * 16 + 2 * t;
        mma16816(dv_acc[ot], a0, a1, a2, a3, lds32(bo), lds32(bo + 8));
        mma16816(dk_acc[ot], e0, e1, e2, e3, lds32(bq), lds32(bq + 8));
      }
    }
  }

  // ---- dK / dV slices of this warp's 16 keys -> global (atomic: several query rows may share a partner row)
  const size_t krow0 = (size_t)rr * N + (size_t)jb * BLK + warp * 16 + g;
#pragma unroll
  for (int ot = 0; ot < OT; ++ot) {
    const size_t col = (size_t)h * D + c0 + ot * 8 + 2 * t;
    float* k0 = dk + krow0 * C + col;
    float* k1 = dk + (krow0 + 8) * C + col;
    float* v0 = dv + krow0 * C + col;
    float* v1 = dv + (krow0 + 8) * C + col;
    red_add2(k0, dk_acc[ot][0] * scale, dk_acc[ot][1] * scale);
    red_add2(k1, dk_acc[ot][2] * scale, dk_acc[ot][3] * scale);
    red_add2(v0, dv_acc[ot][0], dv_acc[ot][1]);
    red_add2(v1, dv_acc[ot][2], dv_acc[ot][3]);
  }
}

template <int D, int BLK>
static int launch(const void* q, const void* k, const void* v, const float* rowp, const int* partner, const void* d_o,
                  const float* lse, const float* delta, float* dq, float* dk, float* dv, int R, int N, int H, int feat, float S,
                  cudaStream_t st) {
  using Cf = CfgB<D, BLK>;
  static bool configured = false;
  if (!configured) {
    CVD_CUDA(cudaFuncSetAttribute(epi_attn_bwd_kernel<D, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cf::kSmem));
    configured = true;
  }
  dim3 grid((unsigned)((N / BLK) * Cf::kChunks), (unsigned)H, (unsigned)R);
  epi_attn_bwd_kernel<D, BLK><<<grid, Cf::kWarps * 32, Cf::kSmem, st>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, rowp, partner, (const __nv_bfloat16*)d_o, lse,
      delta, dq, dk, dv, N, H, feat, S / (float)feat);
  CVD_LAUNCH_CHECK("epi_attn_bwd_kernel");
  return CVD_OK;
}

}  // namespace bwd
}  // namespace cvd

extern "C" int cvd_epi_attn_bwd(const void* q, const void* k, const void* v, const float* rowp, const int* partner,
                                const void* out, const void* d_out, const float* lse, float* delta, float* dq, float* dk,
                                float* dv, int R, int R_kv, int N, int heads, int head_dim, int feat, float F_mat_size,
                                cvd_stream_t stream) {
  using namespace cvd;
  (void)R_kv;   // rows of k / v / dk / dv: bounds the partner indices, which the caller guarantees
  CVD_REQUIRE(q && k && v && partner && out && d_out && lse && delta && dq && dk && dv, "cvd_epi_attn_bwd: null pointer");
  CVD_REQUIRE(R > 0 && R <= 65535 && heads > 0 && heads <= 65535, "cvd_epi_attn_bwd: R=%d heads=%d", R, heads);
  CVD_REQUIRE(N == feat * feat, "cvd_epi_attn_bwd: N=%d is not feat^2 (feat=%d)", N, feat);
  CVD_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)d_out) & 15) == 0,
              "cvd_epi_attn_bwd: q/k/v/out/d_out must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = (long long)R * N * heads;
  bwd::attn_bwd_delta_kernel<<<(unsigned)((items + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)d_out,
                                                                        delta, items, N, heads, head_dim);
  CVD_LAUNCH_CHECK("attn_bwd_delta_kernel");
#define CVD_BWD(D_) \
  if (head_dim == D_) { \
    if (N % 64 == 0) return bwd::launch<D_, 64>(q, k, v, rowp, partner, d_out, lse, delta, dq, dk, dv, R, N, heads, feat, F_mat_size, st); \
    if (N % 16 == 0) return bwd::launch<D_, 16>(q, k, v, rowp, partner, d_out, lse, delta, dq, dk, dv, R, N, heads, feat, F_mat_size, st); \
  }
  CVD_BWD(40)
  CVD_BWD(80)
  CVD_BWD(160)
#undef CVD_BWD
  return fail(CVD_E_INVALID, "cvd_epi_attn_bwd: (head_dim=%d, N=%d) is not an instantiated configuration", head_dim, N);
}
