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
// ldmatrix: four 8x8 bf16 tiles per instruction; lane l supplies the address of row (l & 7) of tile (l >> 3).
// .trans hands every thread the element pair of the transposed tile.
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
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
  // Q / dO tiles double-buffered (cp.async) where two CTAs still fit per SM; d = 160 (N <= 64 in the UNet: one or
  // four query blocks) keeps one stage and two CTAs per SM instead
  static constexpr int kQStages = D <= 80 ? 2 : 1;
  static constexpr size_t kSmem = (size_t)((2 + 2 * kQStages) * BLK * kDP + 2 * BLK * kTP) * sizeof(__nv_bfloat16);
};

// rows [row0, row0 + BLK) of a [rows, C] bf16 matrix, columns [h*D, h*D + D) -> dst [BLK][DP] (pad columns zero)
template <int D, int BLK>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ src, size_t row0, int C, int h, __nv_bfloat16* dst) {
  using Cf = CfgB<D, BLK>;
  constexpr int kVec = Cf::kD16 / 8;
  for (int i = threadIdx.x; i < BLK * kVec; i += Cf::kWarps * 32) {
    const int row = i / kVec, v = i % kVec;
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (v * 8 < D) t = *reinterpret_cast<const uint4*>(src + (row0 + row) * (size_t)C + (size_t)h * D + v * 8);
    *reinterpret_cast<uint4*>(dst + row * Cf::kDP + v * 8) = t;
  }
}

// same tile through cp.async (16-byte chunks, pad columns zero-filled by a zero source size): the copy of the next query
// block is in flight while the current one is multiplied
template <int D, int BLK>
__device__ __forceinline__ void load_tile_async(const __nv_bfloat16* __restrict__ src, size_t row0, int C, int h, __nv_bfloat16* dst) {
  using Cf = CfgB<D, BLK>;
  constexpr int kVec = Cf::kD16 / 8;
  for (int i = threadIdx.x; i < BLK * kVec; i += Cf::kWarps * 32) {
    const int row = i / kVec, v = i % kVec;
    const bool in = v * 8 < D;
    const __nv_bfloat16* g = src + (row0 + row) * (size_t)C + (size_t)h * D + (in ? v * 8 : 0);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + row * Cf::kDP + v * 8);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(in ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

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
  constexpr int QS = Cf::kQStages;
  __nv_bfloat16* sQ2 = reinterpret_cast<__nv_bfloat16*>(bwd_smem);   // [stages][BLK][DP]
  __nv_bfloat16* sdO2 = sQ2 + QS * BLK * DP;                         // [stages][BLK][DP]
  __nv_bfloat16* sK = sdO2 + QS * BLK * DP;
  __nv_bfloat16* sV = sK + BLK * DP;
  __nv_bfloat16* sP = sV + BLK * DP;           // [BLK queries][TP]  P   (row-major; transposed on the way out by ldmatrix.trans)
  __nv_bfloat16* sdS = sP + BLK * TP;          // [BLK queries][TP]  dS

  const int C = H * D;
  const int jb = blockIdx.x / Cf::kChunks, c0 = (blockIdx.x % Cf::kChunks) * kDH;   // key block, head-dim slice
  const int h = blockIdx.y, r = blockIdx.z;
  const int rr = partner[r];                     // K / V row of this query row (single partner)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf((float)D), scale2 = scale * kLog2e;
  const float off = (step - 1.0f) * 0.5f;

  const int lr = lane & 7, lm = lane >> 3;       // ldmatrix: row and tile this lane addresses
  load_tile<D, BLK>(k, (size_t)rr * N + (size_t)jb * BLK, C, h, sK);
  load_tile<D, BLK>(v, (size_t)rr * N + (size_t)jb * BLK, C, h, sV);

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
  if (QS == 2) {
    load_tile_async<D, BLK>(q, (size_t)r * N, C, h, sQ2);
    load_tile_async<D, BLK>(d_o, (size_t)r * N, C, h, sdO2);
    cp_async_commit();
  }
  for (int ib = 0; ib < nqb; ++ib) {
    __syncthreads();                             // previous iteration's readers of the other stage / sP / sdS are done
    const int stg = QS == 2 ? (ib & 1) : 0;
    const __nv_bfloat16* sQ = sQ2 + stg * BLK * DP;
    const __nv_bfloat16* sdO = sdO2 + stg * BLK * DP;
    if (QS == 1) {
      load_tile<D, BLK>(q, (size_t)r * N + (size_t)ib * BLK, C, h, sQ2);
      load_tile<D, BLK>(d_o, (size_t)r * N + (size_t)ib * BLK, C, h, sdO2);
    } else if (ib + 1 < nqb) {                   // next query block: in flight under this block's MMAs
      load_tile_async<D, BLK>(q, (size_t)r * N + (size_t)(ib + 1) * BLK, C, h, sQ2 + ((ib + 1) & 1) * BLK * DP);
      load_tile_async<D, BLK>(d_o, (size_t)r * N + (size_t)(ib + 1) * BLK, C, h, sdO2 + ((ib + 1) & 1) * BLK * DP);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- S = Q K^T and dP = dO V^T for this warp's 16 query rows x BLK keys
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t aq[4], ad[4];
      ldsm4(aq, sQ + (warp * 16 + lr + 8 * (lm & 1)) * DP + ks * 16 + 8 * (lm >> 1));
      ldsm4(ad, sdO + (warp * 16 + lr + 8 * (lm & 1)) * DP + ks * 16 + 8 * (lm >> 1));
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {       // two key tiles per ldmatrix.x4
        uint32_t bk[4], bv[4];
        ldsm4(bk, sK + ((nt + (lm >> 1)) * 8 + lr) * DP + ks * 16 + 8 * (lm & 1));
        ldsm4(bv, sV + ((nt + (lm >> 1)) * 8 + lr) * DP + ks * 16 + 8 * (lm & 1));
        mma16816(s[nt], aq[0], aq[1], aq[2], aq[3], bk[0], bk[1]);
        mma16816(s[nt + 1], aq[0], aq[1], aq[2], aq[3], bk[2], bk[3]);
        mma16816(dp[nt], ad[0], ad[1], ad[2], ad[3], bv[0], bv[1]);
        mma16816(dp[nt + 1], ad[0], ad[1], ad[2], ad[3], bv[2], bv[3]);
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
      float pv[4];
      // this thread's two keys of the tile, key0 and key0 + 1, sit in one key-grid row (feat is even): their grid
      // coordinates once per tile, not per element (the per-element integer division was most of the instruction stream)
      const int key0 = jb * BLK + nt * 8 + 2 * t;
      const float colf = (float)(key0 % feat), rowf = (float)(key0 / feat);
      float tt0[2], tt1[2];                              // line distance terms of rows g / g + 8 at key0 / key0 + 1
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        tt0[i] = fmaf(As[i], colf, fmaf(Bs[i], rowf, base[i]));
        tt1[i] = tt0[i] + As[i];
      }
      // a non-finite row carries an all-zero table (cvd_epi_row_params, epi_module.py:404-407): bias = 0 without a
      // per-element check, as in the forward kernels.  The two keys of a thread are one packed fp32x2 lane pair:
      //   x = S * scale2 + min(band' - l - |t|, -l),   P = 2^x,   dS = P * (dP - delta)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float bm = bandp - l2[i], nl = -l2[i];
        const float v0 = rowp != nullptr ? fminf(bm - fabsf(tt0[i]), nl) : nl;
        const float v1 = rowp != nullptr ? fminf(bm - fabsf(tt1[i]), nl) : nl;
        float x0, x1;
        ptx::unpack2(ptx::fma2(ptx::pack2(s[nt][2 * i], s[nt][2 * i + 1]), ptx::pack2(scale2, scale2), ptx::pack2(v0, v1)), x0, x1);
        pv[2 * i] = ptx::ex2(x0);
        pv[2 * i + 1] = ptx::ex2(x1);
        const uint64_t dd = ptx::add2(ptx::pack2(dp[nt][2 * i], dp[nt][2 * i + 1]), ptx::pack2(-dl[i], -dl[i]));
        ptx::unpack2(ptx::fma2(ptx::pack2(pv[2 * i], pv[2 * i + 1]), dd, ptx::pack2(0.f, 0.f)), s[nt][2 * i], s[nt][2 * i + 1]);
      }
      // row-major [query][key], one 32-bit store per key pair (conflict-free: 36-word pitch)
      __nv_bfloat16* p0 = sP + (warp * 16 + g) * TP + nt * 8 + 2 * t;
      __nv_bfloat16* d0 = sdS + (warp * 16 + g) * TP + nt * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(p0) = packbf(pv[0], pv[1]);
      *reinterpret_cast<uint32_t*>(p0 + 8 * TP) = packbf(pv[2], pv[3]);
      *reinterpret_cast<uint32_t*>(d0) = packbf(s[nt][0], s[nt][1]);
      *reinterpret_cast<uint32_t*>(d0 + 8 * TP) = packbf(s[nt][2], s[nt][3]);
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
        for (int ot = 0; ot < OT; ot += 2) {     // K[key][d] as the (key x d) B operand: ldmatrix.trans, two d tiles at a time
          const int o2 = ot + 1 < OT ? ot + (lm >> 1) : ot;
          uint32_t bk[4];
          ldsm4t(bk, sK + (kk * 16 + lr + 8 * (lm & 1)) * DP + c0 + o2 * 8);
          mma16816(dqa[ot], a0, a1, a2, a3, bk[0], bk[1]);
          if (ot + 1 < OT) mma16816(dqa[ot + 1], a0, a1, a2, a3, bk[2], bk[3]);
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
#pragma unroll
    for (int kk = 0; kk < BLK / 16; ++kk) {
      // A = P^T / dS^T (keys x queries) out of the row-major [query][key] tiles: ldmatrix.trans
      uint32_t ap[4], as_[4];
      ldsm4t(ap, sP + (kk * 16 + lr + 8 * (lm >> 1)) * TP + warp * 16 + 8 * (lm & 1));
      ldsm4t(as_, sdS + (kk * 16 + lr + 8 * (lm >> 1)) * TP + warp * 16 + 8 * (lm & 1));
#pragma unroll
      for (int ot = 0; ot < OT; ot += 2) {       // B = dO / Q (queries x d): ldmatrix.trans, two d tiles at a time
        const int o2 = ot + 1 < OT ? ot + (lm >> 1) : ot;
        uint32_t bo[4], bq[4];
        ldsm4t(bo, sdO + (kk * 16 + lr + 8 * (lm & 1)) * DP + c0 + o2 * 8);
        ldsm4t(bq, sQ + (kk * 16 + lr + 8 * (lm & 1)) * DP + c0 + o2 * 8);
        mma16816(dv_acc[ot], ap[0], ap[1], ap[2], ap[3], bo[0], bo[1]);
        mma16816(dk_acc[ot], as_[0], as_[1], as_[2], as_[3], bq[0], bq[1]);
        if (ot + 1 < OT) {
          mma16816(dv_acc[ot + 1], ap[0], ap[1], ap[2], ap[3], bo[2], bo[3]);
          mma16816(dk_acc[ot + 1], as_[0], as_[1], as_[2], as_[3], bq[2], bq[3]);
        }
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
  CVD_CUDA(ensure_dynamic_smem((const void*)epi_attn_bwd_kernel<D, BLK>, Cf::kSmem));
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
  // refuse before the first launch (the delta kernel below is shape-agnostic)
  if (!((head_dim == 40 || head_dim == 80 || head_dim == 160) && N % 16 == 0))
    return fail(CVD_E_INVALID, "cvd_epi_attn_bwd: (head_dim=%d, N=%d) is not an instantiated configuration", head_dim, N);
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
