// Fused epipolar cross-view attention forward for sm_100a (B200):
//     out = softmax_k( q k^T / sqrt(d) + bias[r, q, k] ) v
// replacing get_attn_map + prepare_attention_mask + F.scaled_dot_product_attention of the
// reference (animatediff/models/epi_module.py:293-329,395-407;
// animatediff/models/attention_processor.py:573-619).
//
//  * The [R,N,N] bias is never read from memory: each softmax thread owns one query pixel,
//    derives its epipolar line l_q = M_r x_q from the 12-float row table once, and evaluates
//    bias = min(band' - |A x_k + B y_k + C|, 0) in registers (base-2 units, 3 FP32 ops/key).
//  * S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma, cta_group::1,
//    M=128): Q/K/V^T tiles arrive by TMA (128B swizzle) into shared memory, S and O live in
//    TMEM, P goes back to TMEM as the A operand of the second MMA (never through shared
//    memory), and the softmax denominator is a column of O (V^T carries a row of ones).
//  * Warp roles per CTA (192 threads): warps 0-3 softmax (one TMEM lane quarter each),
//    warp 4 TMA producer, warp 5 MMA issuer + TMEM owner.  All hand-offs are mbarriers;
//    S is double-buffered so the next Q K^T overlaps the current softmax.
//  * One CTA per (row, head, 128-query block); two CTAs are resident per SM for d=40/80 so
//    that one CTA's exp/bias math covers the other's barrier latencies.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cvd {
namespace sm100 {

using namespace ptx;

constexpr int kBlockM = 128;   // queries per CTA (= TMEM lanes)
constexpr int kBlockN = 64;    // keys per tile
constexpr int kThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;   // lazy O rescale: only when the row max grows by > 2^8

template <int D>
struct Cfg {
  static_assert(D % 8 == 0 && (D % 16 == 0 || D % 16 == 8), "head_dim must be a multiple of 8");
  static constexpr int kKBlocks = (D + 63) / 64;          // 128-byte swizzle atoms along the head dim
  static constexpr int kKSteps = (D + 15) / 16;           // UMMA K steps of Q K^T
  static constexpr int kDV = (D / 16 + 1) * 16;           // rows of a V^T tile: d values, a row of ones, zero pad
  static constexpr int kQBytes = kKBlocks * kBlockM * 128;
  static constexpr int kKBytes = kKBlocks * kBlockN * 128;
  static constexpr int kVBytes = kDV * 128;
  static constexpr int kKStages = D <= 40 ? 4 : 3;
  static constexpr int kVStages = D <= 40 ? 4 : 2;
  // TMEM columns: S0 | S1 | P | O
  static constexpr int kColS = 0, kColP = 2 * kBlockN, kColO = kColP + kBlockN / 2;
  static constexpr int kTmemCols = (kColO + kDV) <= 256 ? 256 : 512;
  static constexpr int kMinBlocks = (kTmemCols == 256) ? 2 : 1;
  static constexpr int kSmemBytes = kQBytes + kKStages * kKBytes + kVStages * kVBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Barriers {
  uint64_t q_full, q_ready, p_full, o_done;
  uint64_t s_full[2], s_free[2];
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint32_t tmem_base;
};

template <int D, int FEAT, bool DBG>
__global__ void __launch_bounds__(kThreads, Cfg<D>::kMinBlocks)
epi_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const float* __restrict__ rowp,
                    const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, int R, int H, int n_partners,
                    float step, float* __restrict__ dbg) {
  using C = Cfg<D>;
  constexpr int N = FEAT * FEAT;
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  constexpr int kTilesPerPartner = (N + kBlockN - 1) / kBlockN;
  constexpr int kRowsPerTile = (kBlockN + FEAT - 1) / FEAT;   // key-grid rows touched by one key tile
  constexpr uint32_t kIdescQK = umma_idesc_bf16(kBlockM, kBlockN);
  constexpr uint32_t kIdescPV = umma_idesc_bf16(kBlockM, C::kDV);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + C::kQBytes;
  const uint32_t sV = sK + C::kKStages * C::kKBytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_raw + (sV + C::kVStages * C::kVBytes - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % kQTiles;
  const int h = (blockIdx.x / kQTiles) % H;
  const int r = blockIdx.x / (kQTiles * H);
  const int total = n_partners * kTilesPerPartner;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(smem_u32(&bars->q_full), 1);
    mbar_init(smem_u32(&bars->q_ready), 128);
    mbar_init(smem_u32(&bars->p_full), 128);
    mbar_init(smem_u32(&bars->o_done), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_free[i]), 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<C::kTmemCols>(smem_u32(&bars->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::kQBytes);
#pragma unroll
      for (int kb = 0; kb < C::kKBlocks; ++kb)
        tma_load_2d(sQ + kb * kBlockM * 128, &tmQ, h * D + kb * 64, r * N + qt * kBlockM, smem_u32(&bars->q_full));
      for (int it = 0; it <= total; ++it) {
        if (it < total) {
          const int p = it / kTilesPerPartner, j = it % kTilesPerPartner;
          const int rr = partner[p * R + r];
          const int st = it % C::kKStages;
          if (it >= C::kKStages) mbar_wait(smem_u32(&bars->k_empty[st]), ((it / C::kKStages) - 1) & 1);
          mbar_expect_tx(smem_u32(&bars->k_full[st]), C::kKBytes);
#pragma unroll
          for (int kb = 0; kb < C::kKBlocks; ++kb)
            tma_load_2d(sK + st * C::kKBytes + kb * kBlockN * 128, &tmK, h * D + kb * 64, rr * N + j * kBlockN,
                        smem_u32(&bars->k_full[st]));
        }
        if (it >= 1) {
          const int jt = it - 1;
          const int p = jt / kTilesPerPartner, j = jt % kTilesPerPartner;
          const int rr = partner[p * R + r];
          const int st = jt % C::kVStages;
          if (jt >= C::kVStages) mbar_wait(smem_u32(&bars->v_empty[st]), ((jt / C::kVStages) - 1) & 1);
          mbar_expect_tx(smem_u32(&bars->v_full[st]), C::kVBytes);
          tma_load_2d(sV + st * C::kVBytes, &tmV, j * kBlockN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]));
        }
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      mbar_wait(smem_u32(&bars->q_ready), 0);
      tc_fence_after();
      for (int it = 0; it <= total; ++it) {
        if (it < total) {   // S[it&1] = Q K_it^T
          const int st = it % C::kKStages, buf = it & 1;
          mbar_wait(smem_u32(&bars->k_full[st]), (it / C::kKStages) & 1);
          if (it >= 2) mbar_wait(smem_u32(&bars->s_free[buf]), ((it >> 1) - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < C::kKSteps; ++ks) {
            const uint32_t koff = (ks >> 2) * kBlockM * 128 + (ks & 3) * 32;
            const uint32_t koffk = (ks >> 2) * kBlockN * 128 + (ks & 3) * 32;
            umma_ss(tmem + C::kColS + buf * kBlockN, umma_desc_k_sw128(sQ + koff),
                    umma_desc_k_sw128(sK + st * C::kKBytes + koffk), kIdescQK, ks > 0);
          }
          umma_commit(smem_u32(&bars->s_full[buf]));
          umma_commit(smem_u32(&bars->k_empty[st]));
        }
        if (it >= 1) {      // O += P_jt V_jt
          const int jt = it - 1, st = jt % C::kVStages;
          mbar_wait(smem_u32(&bars->v_full[st]), (jt / C::kVStages) & 1);
          mbar_wait(smem_u32(&bars->p_full), jt & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kBlockN / 16; ++ks)
            umma_ts(tmem + C::kColO, tmem + C::kColP + ks * 8, umma_desc_k_sw128(sV + st * C::kVBytes + ks * 32),
                    kIdescPV, (jt > 0) || (ks > 0));
          umma_commit(smem_u32(&bars->v_empty[st]));
          umma_commit(smem_u32(&bars->o_done));
        }
      }
    }
  } else {
    // =============================== softmax warps ==============================
    const int tid = threadIdx.x;                       // query row inside the tile == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    int n = qt * kBlockM + tid;
    const bool row_valid = n < N;
    if (!row_valid) n = N - 1;
    const float xq = pixel_centre(n % FEAT, step), yq = pixel_centre(n / FEAT, step);
    const float off = (step - 1.0f) * 0.5f;
    const float scale_log2 = kLog2e * rsqrtf((float)D);

    mbar_wait(smem_u32(&bars->q_full), 0);
    if (D % 16 != 0) {
      // zero the K-padding columns [D, D+8) of this thread's Q row (16-byte chunk D/8 of the
      // last swizzle atom, stored at chunk ^ (row & 7)) so that K's neighbouring head adds 0
      constexpr int kb = D / 64, chunk = (D % 64) / 8;
      unsigned char* qrow = smem_raw + (sQ - smem_u32(smem_raw)) + kb * kBlockM * 128 + tid * 128 + ((chunk ^ (tid & 7)) << 4);
      *reinterpret_cast<uint4*>(qrow) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    mbar_arrive(smem_u32(&bars->q_ready));

    // bias geometry of this thread's query pixel, base-2 units (already scaled by w_r * log2e / |l_q|):
    //   t(key) = As * col + Bs * row + base,   bias = min(band' - |t|, 0) = min(band' - t, band' + t, 0)
    float As = 0.f, Bs = 0.f, base = 0.f, bandp = 0.f;
    float m_used = 0.f;                   // running reference maximum (0 until the first tile sets it)
    uint32_t sreg[kBlockN];
    uint32_t preg[kBlockN / 2];
    const uint64_t scale2 = pack2(scale_log2, scale_log2);

    for (int it = 0; it < total; ++it) {
      const int p = it / kTilesPerPartner, j = it % kTilesPerPartner;
      if (j == 0) {
        if (rowp != nullptr) {
          const float* P = rowp + (size_t)(p * R + r) * CVD_ROWP_STRIDE;
          const float a = fmaf(P[1], yq, P[0] * xq) + P[2];
          const float b = fmaf(P[4], yq, P[3] * xq) + P[5];
          const float c = fmaf(P[7], yq, P[6] * xq) + P[8];
          const float k1 = P[10] * kLog2e / (sqrtf(a * a + b * b) + 1e-6f);
          const float A = a * k1, B = b * k1;
          As = A * step;
          Bs = B * step;
          base = fmaf(A, off, fmaf(B, off, c * k1));
          bandp = P[9] * P[10] * kLog2e;
        }
      }
      const int buf = it & 1;
      // DBG build: clock64 stamps of thread 0 at the hand-off points of every iteration, written
      // behind the S dump: int64 [grid, 32, 8]
      long long* stamps = nullptr;
      if (DBG && tid == 0 && it < 32)
        stamps = reinterpret_cast<long long*>(dbg + (size_t)gridDim.x * kBlockM * kBlockN) + ((size_t)blockIdx.x * 32 + it) * 8;
      if (DBG && stamps) stamps[0] = clock64();
      mbar_wait(smem_u32(&bars->s_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      if (DBG && stamps) stamps[1] = clock64();
      tmem_ld_x32(tlane + C::kColS + buf * kBlockN, sreg);
      tmem_ld_x32(tlane + C::kColS + buf * kBlockN + 32, sreg + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_free[buf]));
      if (DBG && stamps) stamps[2] = clock64();

      if (DBG && it == 0) {
        float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kBlockN;
#pragma unroll
        for (int c = 0; c < kBlockN; ++c) o[c] = __uint_as_float(sreg[c]);
      }

      // x = qk * log2e/sqrt(d) + bias - m_used, evaluated against the CURRENT reference maximum so
      // the subtraction is free: min(band'-m - t, band'-m + t, -m) is one FMNMX3, the two linear
      // terms advance by one packed FADD2 per key pair, and the score is one packed FFMA2.
      const float bm = bandp - m_used, negm = -m_used;
      const uint64_t step1 = pack2(-2.f * As, -2.f * As), step2 = pack2(2.f * As, 2.f * As);
      float xmax = -INFINITY;
      uint64_t a1 = 0, a2 = 0;
#pragma unroll
      for (int c = 0; c < kBlockN; c += 2) {
        if (c % FEAT == 0) {   // start of a key-grid row inside the tile
          const float rb = fmaf(Bs, (float)(j * kRowsPerTile + c / FEAT), base);
          a1 = pack2(bm - rb, bm - rb - As);
          a2 = pack2(bm + rb, bm + rb + As);
        }
        float a1x, a1y, a2x, a2y;
        unpack2(a1, a1x, a1y);
        unpack2(a2, a2x, a2y);
        const float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
        float x0, x1;
        unpack2(fma2(pack2(__uint_as_float(sreg[c]), __uint_as_float(sreg[c + 1])), scale2, pack2(v0, v1)), x0, x1);
        if (N < kBlockN && c >= N) x0 = x1 = -INFINITY;   // keys past the row (N = 16): not part of the softmax
        xmax = max3(xmax, x0, x1);
        sreg[c] = __float_as_uint(x0);
        sreg[c + 1] = __float_as_uint(x1);
        a1 = add2(a1, step1);
        a2 = add2(a2, step2);
      }

      if (DBG && stamps) stamps[3] = clock64() + (long long)(__float_as_uint(xmax) & 1u);
      // lazy maximum update: only when this tile exceeds the reference by more than 2^8 (or on the
      // very first tile) is the reference moved, O rescaled and x shifted.
      const bool need = (it == 0) || (xmax > kRescaleThreshold);
      if (__any_sync(0xffffffffu, need)) {
        if (it > 0) {
          mbar_wait(smem_u32(&bars->o_done), (it - 1) & 1);   // PV(it-1) must have landed in O
          tc_fence_after();
          const float f = need ? ex2(-xmax) : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < C::kDV; c0 += 16) {
            uint32_t o[16];
            tmem_ld_x16(tlane + C::kColO + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_x16(tlane + C::kColO + c0, o);
          }
          tmem_st_wait();
        }
        if (need) {
          m_used += xmax;
#pragma unroll
          for (int c = 0; c < kBlockN; ++c) sreg[c] = __float_as_uint(__uint_as_float(sreg[c]) - xmax);
        }
      }
#pragma unroll
      for (int c = 0; c < kBlockN; c += 2)
        preg[c >> 1] = pack_bf16(ex2(__uint_as_float(sreg[c])), ex2(__uint_as_float(sreg[c + 1])));
      if (DBG && stamps) stamps[4] = clock64() + (long long)(preg[0] & 1u);
      if (it > 0) {   // PV(it-1) has finished reading the single P buffer
        mbar_wait(smem_u32(&bars->o_done), (it - 1) & 1);
        tc_fence_after();
      }
      if (DBG && stamps) stamps[5] = clock64();
      tmem_st_x32(tlane + C::kColP, preg);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full));
      if (DBG && stamps) stamps[6] = clock64();
    }

    // epilogue: O[:, 0:D] / O[:, D]  ->  out[r, n, h*D : (h+1)*D]
    mbar_wait(smem_u32(&bars->o_done), (total - 1) & 1);
    tc_fence_after();
    uint32_t lraw;
    tmem_ld_x1(tlane + C::kColO + D, lraw);
    tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(lraw);
    __nv_bfloat16* orow = out + ((size_t)r * N + n) * (size_t)(H * D) + h * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 8) {
      uint32_t o[8];
      tmem_ld_x8(tlane + C::kColO + c0, o);
      tmem_ld_wait();
      uint4 pk;
      pk.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      pk.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      pk.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      pk.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      if (row_valid) *reinterpret_cast<uint4*>(orow + c0) = pk;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous), box [box_rows, 64 cols], 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(CVD_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    return fail(CVD_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)rc,
                (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return CVD_OK;
}

template <int D, int FEAT>
static int launch(const void* q, const void* k, const void* vt, const float* rowp, const int* partner, void* out, int R,
                  int R_kv, int H, int n_partners, float S, float* dbg, cudaStream_t st) {
  using C = Cfg<D>;
  constexpr int N = FEAT * FEAT;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_map(&tmQ, q, (uint64_t)R * N, (uint64_t)H * D, kBlockM)) != CVD_OK) return rc;
  if ((rc = make_map(&tmK, k, (uint64_t)R_kv * N, (uint64_t)H * D, kBlockN)) != CVD_OK) return rc;
  if ((rc = make_map(&tmV, vt, (uint64_t)R_kv * H * C::kDV, (uint64_t)N, C::kDV)) != CVD_OK) return rc;
  constexpr bool kHasDebug = (D == 40 && FEAT == 32);   // the S-dump variant is built for the bench shape only
  CVD_REQUIRE(dbg == nullptr || kHasDebug, "cvd_epi_attn_fwd_debug: only (head_dim=40, feat=32) has a debug build");
  auto kern = (dbg != nullptr && kHasDebug) ? epi_attn_fwd_kernel<D, FEAT, kHasDebug> : epi_attn_fwd_kernel<D, FEAT, false>;
  CVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  const long long grid = (long long)R * H * kQTiles;
  CVD_REQUIRE(grid > 0 && grid < (1ll << 31), "cvd_epi_attn_fwd: grid of %lld CTAs", grid);
  kern<<<(unsigned)grid, kThreads, C::kSmemBytes, st>>>(tmQ, tmK, tmV, rowp, partner, (__nv_bfloat16*)out, R, H, n_partners,
                                                        S / (float)FEAT, dbg);
  CVD_LAUNCH_CHECK("epi_attn_fwd_kernel");
  return CVD_OK;
}

static int dispatch(const void* q, const void* k, const void* vt, const float* rowp, const int* partner, void* out, int R,
                    int R_kv, int N, int heads, int head_dim, int n_partners, int feat, float S, float* dbg, cudaStream_t st) {
  CVD_REQUIRE(q && k && vt && partner && out, "cvd_epi_attn_fwd: null pointer");
  CVD_REQUIRE(R > 0 && R_kv > 0 && heads > 0 && n_partners >= 1, "cvd_epi_attn_fwd: R=%d R_kv=%d heads=%d n_partners=%d", R, R_kv, heads,
              n_partners);
  CVD_REQUIRE(N == feat * feat, "cvd_epi_attn_fwd: N=%d is not feat^2 (feat=%d); only square latents are supported", N, feat);
  CVD_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vt | (uintptr_t)out) & 15) == 0,
              "cvd_epi_attn_fwd: q/k/vt/out must be 16-byte aligned");
#define CVD_CASE(D_, F_) \
  if (head_dim == D_ && feat == F_) return launch<D_, F_>(q, k, vt, rowp, partner, out, R, R_kv, heads, n_partners, S, dbg, st);
  CVD_CASE(40, 32)
  CVD_CASE(40, 64)
  CVD_CASE(80, 16)
  CVD_CASE(80, 32)
  CVD_CASE(160, 8)
  CVD_CASE(160, 16)
  CVD_CASE(160, 4)
  CVD_CASE(40, 16)
  CVD_CASE(40, 8)
  CVD_CASE(40, 4)
  CVD_CASE(80, 8)
  CVD_CASE(80, 4)
#undef CVD_CASE
  return fail(CVD_E_INVALID, "cvd_epi_attn_fwd: (head_dim=%d, feat=%d) is not an instantiated configuration", head_dim, feat);
}

}  // namespace sm100
}  // namespace cvd

extern "C" int cvd_epi_attn_fwd(const void* q, const void* k, const void* vt, const float* rowp, const int* partner,
                                void* out, int R, int R_kv, int N, int heads, int head_dim, int n_partners, int feat,
                                float F_mat_size, cvd_stream_t stream) {
  return cvd::sm100::dispatch(q, k, vt, rowp, partner, out, R, R_kv, N, heads, head_dim, n_partners, feat, F_mat_size, nullptr,
                              (cudaStream_t)stream);
}

// Bring-up aid: additionally dumps the raw Q K^T accumulator of key tile 0 of every CTA
// (dbg_s: fp32 [grid, 128, 64]) so the first MMA can be verified independently of the rest.
extern "C" int cvd_epi_attn_fwd_debug(const void* q, const void* k, const void* vt, const float* rowp, const int* partner,
                                      void* out, float* dbg_s, int R, int R_kv, int N, int heads, int head_dim,
                                      int n_partners, int feat, float F_mat_size, cvd_stream_t stream) {
  return cvd::sm100::dispatch(q, k, vt, rowp, partner, out, R, R_kv, N, heads, head_dim, n_partners, feat, F_mat_size, dbg_s,
                              (cudaStream_t)stream);
}
