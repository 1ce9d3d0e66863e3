// Fused epipolar cross-view attention forward for sm_100a (B200):
//     out = softmax_k( q k^T / sqrt(d) + bias[r, q, k] ) v
// replacing get_attn_map + prepare_attention_mask + F.scaled_dot_product_attention of the
// reference (animatediff/models/epi_module.py:293-329,395-407;
// animatediff/models/attention_processor.py:573-619).
//
//  * The [R,N,N] bias is never read from memory: each softmax thread owns one query pixel,
//    derives its epipolar line l_q = M_r x_q from the 12-float row table once, and evaluates
//    bias = min(band' - t, band' + t, 0) in registers (base-2 units): per key pair two packed
//    FADD2 advance the two linear terms, one FMNMX3 per key takes the minimum, one packed FFMA2
//    adds the scaled score.
//  * S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma, cta_group::1,
//    M=128): Q/K/V^T tiles arrive by TMA (128B swizzle) into shared memory, S and O live in
//    TMEM, P goes back to TMEM *over the S buffer it came from* as the A operand of the second
//    MMA (never through shared memory), and the softmax denominator is a column of O (V^T
//    carries a row of ones).
//  * The kernel is bound by the exp unit (MUFU: 16 ex2/clk/SM) and by hand-off latencies, not
//    by the MMA (at d=40: 2 x 128x32x48 MACs per 4096 exps).  So the design maximises resident
//    softmax warps: keys are consumed in 32-key sub-tiles, TMEM per CTA is S0|S1 (2 x 32 columns,
//    double-buffered so Q K^T of sub-tile t+2 is issued while the softmax of t runs) plus DV
//    columns of O = 112 <= 128 columns at d=40, a thread holds 32 scores at a time (<= 80
//    registers), and FOUR CTAs are resident per SM (two at d=80/160): one CTA's barrier and
//    TMEM latencies are covered by the other CTAs' exp work.
//  * Warp roles per CTA (192 threads): warps 0-3 softmax (one TMEM lane quarter each),
//    warp 4 TMA producer, warp 5 MMA issuer + TMEM owner.  All hand-offs are mbarriers.
//  * Online softmax with a lazily updated reference maximum: scores are evaluated directly
//    against the current reference (the subtraction is folded into the bias terms); only when
//    a sub-tile exceeds it by more than 2^8 is O rescaled.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cvd {
namespace sm100 {

using namespace ptx;

constexpr int kBlockM = 128;   // queries per CTA (= TMEM lanes)
constexpr int kLoadN = 64;     // keys per TMA tile (one 128-byte swizzle row of V^T)
constexpr int kSubN = 32;      // keys per MMA / softmax sub-tile
constexpr int kThreads = 192;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 16.0f;  // lazy rescale: only when a score exceeds the reference by > 2^16

template <int D, int FEAT>
struct Cfg {
  static_assert(D % 8 == 0 && (D % 16 == 0 || D % 16 == 8), "head_dim must be a multiple of 8");
  static constexpr int kN = FEAT * FEAT;
  static constexpr int kSubTiles = (kN + kSubN - 1) / kSubN;      // per partner
  static constexpr int kLoadTiles = (kSubTiles + 1) / 2;          // per partner
  static constexpr int kKBlocks = (D + 63) / 64;          // 128-byte swizzle atoms along the head dim
  static constexpr int kKSteps = (D + 15) / 16;           // UMMA K steps of Q K^T
  static constexpr int kDV = (D / 16 + 1) * 16;           // rows of a V^T tile: d values, a row of ones, zero pad
  static constexpr int kQBytes = kKBlocks * kBlockM * 128;
  static constexpr int kKBytes = kKBlocks * kLoadN * 128;
  static constexpr int kVBytes = kDV * 128;
  // TMEM columns: S0 | S1 (P of a sub-tile overwrites the first 16 columns of its S buffer) | O
  static constexpr int kColS = 0, kColO = 2 * kSubN;
  static constexpr int kTmemCols = (kColO + kDV) <= 128 ? 128 : 256;
  static_assert(kColO + kDV <= 256, "TMEM budget");
  // K/V ring depth: as deep as the residency target allows, never deeper than the tile count
  static constexpr int kWant = D <= 40 ? 3 : (D <= 80 ? 2 : (kLoadTiles > 1 ? 2 : 1));
  static constexpr int kKStages = kWant < kLoadTiles ? kWant : kLoadTiles;
  static constexpr int kWantV = D <= 40 ? 2 : kKStages;
  static constexpr int kVStages = kWantV < kLoadTiles ? kWantV : kLoadTiles;
  static constexpr int kSmemBytes = kQBytes + kKStages * kKBytes + kVStages * kVBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kMinBlocks = D <= 40 ? 4 : (kSmemBytes <= 110 * 1024 ? 2 : 1);
};

struct Barriers {
  uint64_t q_full, q_ready, pv_done, o_done;
  uint64_t s_full[2], p_full[2];
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint32_t tmem_base;
};

template <int D, int FEAT, bool DBG>
__global__ void __launch_bounds__(kThreads, Cfg<D, FEAT>::kMinBlocks)
epi_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const float* __restrict__ rowp,
                    const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, int R, int H, int n_partners,
                    float step, float* __restrict__ dbg) {
  using C = Cfg<D, FEAT>;
  constexpr int N = C::kN;
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  constexpr uint32_t kIdescQK = umma_idesc_bf16(kBlockM, kSubN);
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
  const int total = n_partners * C::kSubTiles;        // sub-tiles this CTA walks
  const int loads = n_partners * C::kLoadTiles;       // TMA tiles this CTA walks

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(smem_u32(&bars->q_full), 1);
    mbar_init(smem_u32(&bars->q_ready), 128);
    mbar_init(smem_u32(&bars->pv_done), 1);
    mbar_init(smem_u32(&bars->o_done), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->p_full[i]), 128);
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
    // The whole warp walks the loop (waits are warp-wide), one elected lane issues the copies.
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::kQBytes);
#pragma unroll
      for (int kb = 0; kb < C::kKBlocks; ++kb)
        tma_load_2d(sQ + kb * kBlockM * 128, &tmQ, h * D + kb * 64, r * N + qt * kBlockM, smem_u32(&bars->q_full));
    }
    for (int lt = 0; lt <= loads; ++lt) {
      if (lt < loads) {       // K tile lt (runs one tile ahead of V)
        const int p = lt / C::kLoadTiles, j = lt % C::kLoadTiles;
        const int rr = partner[p * R + r];
        const int st = lt % C::kKStages;
        if (lt >= C::kKStages) mbar_wait(smem_u32(&bars->k_empty[st]), ((lt / C::kKStages) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->k_full[st]), C::kKBytes);
#pragma unroll
          for (int kb = 0; kb < C::kKBlocks; ++kb)
            tma_load_2d(sK + st * C::kKBytes + kb * kLoadN * 128, &tmK, h * D + kb * 64, rr * N + j * kLoadN,
                        smem_u32(&bars->k_full[st]));
        }
      }
      if (lt >= 1) {          // V^T tile lt-1
        const int lv = lt - 1;
        const int p = lv / C::kLoadTiles, j = lv % C::kLoadTiles;
        const int rr = partner[p * R + r];
        const int st = lv % C::kVStages;
        if (lv >= C::kVStages) mbar_wait(smem_u32(&bars->v_empty[st]), ((lv / C::kVStages) - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bars->v_full[st]), C::kVBytes);
          tma_load_2d(sV + st * C::kVBytes, &tmV, j * kLoadN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]));
        }
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // =============================== MMA issuer =================================
    // Issue order: QK(0), QK(1), then per sub-tile t: PV(t), QK(t+2).  tcgen05.mma executes in issue
    // order, so QK(t+2) -- which overwrites S[t&1], where P_t lives -- cannot overtake PV(t).
    // The whole warp walks the loop; one elected lane issues the MMAs and commits.
    auto issue_qk = [&](int t) {
      const int s = t % C::kSubTiles;
      const int lt = (t / C::kSubTiles) * C::kLoadTiles + (s >> 1);
      const int st = lt % C::kKStages;
      if ((s & 1) == 0) mbar_wait(smem_u32(&bars->k_full[st]), (lt / C::kKStages) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ++ks) {
          const uint32_t qoff = (ks >> 2) * kBlockM * 128 + (ks & 3) * 32;
          const uint32_t koff = (ks >> 2) * kLoadN * 128 + (s & 1) * (kSubN * 128) + (ks & 3) * 32;
          umma_ss(tmem + C::kColS + (t & 1) * kSubN, umma_desc_k_sw128(sQ + qoff),
                  umma_desc_k_sw128(sK + st * C::kKBytes + koff), kIdescQK, ks > 0);
        }
        umma_commit(smem_u32(&bars->s_full[t & 1]));
        if ((s & 1) == 1 || s == C::kSubTiles - 1) umma_commit(smem_u32(&bars->k_empty[st]));
      }
      __syncwarp();
    };
    mbar_wait(smem_u32(&bars->q_ready), 0);
    issue_qk(0);
    if (total > 1) issue_qk(1);
    for (int t = 0; t < total; ++t) {
      const int s = t % C::kSubTiles;
      const int lt = (t / C::kSubTiles) * C::kLoadTiles + (s >> 1);
      const int st = lt % C::kVStages;
      // DBG build: MMA-warp stamps, int64 [grid,32,8] behind the softmax stamps
      long long* ms = nullptr;
      if (DBG && t < 32 && lane == 0)
        ms = reinterpret_cast<long long*>(dbg + (size_t)gridDim.x * kBlockM * kLoadN) + ((size_t)(gridDim.x + blockIdx.x) * 32 + t) * 8;
      if (DBG && ms) ms[0] = clock64();
      if ((s & 1) == 0) mbar_wait(smem_u32(&bars->v_full[st]), (lt / C::kVStages) & 1);
      mbar_wait(smem_u32(&bars->p_full[t & 1]), (t >> 1) & 1);
      tc_fence_after();
      if (DBG && ms) ms[1] = clock64();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kSubN / 16; ++ks)
          umma_ts(tmem + C::kColO, tmem + C::kColS + (t & 1) * kSubN + ks * 8,
                  umma_desc_k_sw128(sV + st * C::kVBytes + (s & 1) * (kSubN * 2) + ks * 32), kIdescPV, (t > 0) || (ks > 0));
        umma_commit(smem_u32(&bars->pv_done));
        if (t == total - 1) umma_commit(smem_u32(&bars->o_done));
        if ((s & 1) == 1 || s == C::kSubTiles - 1) umma_commit(smem_u32(&bars->v_empty[st]));
      }
      __syncwarp();
      if (DBG && ms) ms[3] = ms[2] = clock64();
      if (t + 2 < total) issue_qk(t + 2);
      if (DBG && ms) ms[4] = clock64();
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
    float m_used = 0.f;                   // running reference maximum (set by the first sub-tile)
    float bias_ub = 0.f;                  // upper bound of this query's bias over ALL keys of the first partner
    const uint64_t scale2 = pack2(scale_log2, scale_log2);

    for (int t = 0; t < total; ++t) {
      const int p = t / C::kSubTiles, s = t % C::kSubTiles;
      if (s == 0 && rowp != nullptr) {
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
        if (t == 0) {
          // |t| is convex, so its minimum over the key grid is bounded below by 0 if the corner values
          // change sign and by the smallest corner value otherwise: bias <= min(band' - tmin, 0).
          const float e = (float)(FEAT - 1);
          const float t00 = base, t10 = fmaf(As, e, base), t01 = fmaf(Bs, e, base), t11 = fmaf(As, e, t01);
          const float lo = fminf(fminf(t00, t10), fminf(t01, t11)), hi = fmaxf(fmaxf(t00, t10), fmaxf(t01, t11));
          const float tmin = (lo <= 0.f && hi >= 0.f) ? 0.f : fminf(fabsf(lo), fabsf(hi));
          bias_ub = fminf(bandp - tmin, 0.f);
        }
      }
      // DBG build: clock64 stamps of thread 0 at the hand-off points, behind the S dump: int64 [grid,32,8]
      long long* stamps = nullptr;
      if (DBG && tid == 0 && t < 32)
        stamps = reinterpret_cast<long long*>(dbg + (size_t)gridDim.x * kBlockM * kLoadN) + ((size_t)blockIdx.x * 32 + t) * 8;
      if (DBG && stamps) stamps[0] = clock64();
      const uint32_t tS = tlane + C::kColS + (t & 1) * kSubN;
      mbar_wait(smem_u32(&bars->s_full[t & 1]), (t >> 1) & 1);   // also implies PV(t-2) has completed (issue order)
      tc_fence_after();
      if (DBG && stamps) stamps[1] = clock64();
      uint32_t sreg[kSubN];
      tmem_ld_x32(tS, sreg);
      tmem_ld_wait();
      if (DBG && stamps) stamps[2] = clock64();
      if (DBG && t < 2) {
        float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kLoadN + t * kSubN;
#pragma unroll
        for (int c = 0; c < kSubN; ++c) o[c] = __uint_as_float(sreg[c]);
      }
      // x = qk * log2e/sqrt(d) + bias - m_used, evaluated against the CURRENT reference maximum so
      // the subtraction is free: min(band'-m - t, band'-m + t, -m) is one FMNMX3 per key, the two
      // linear terms advance by one packed FADD2 per key pair, the score is one packed FFMA2.
      const float bm = bandp - m_used, negm = -m_used;
      const uint64_t step1 = pack2(-2.f * As, -2.f * As), step2 = pack2(2.f * As, 2.f * As);
      const int key0 = s * kSubN;          // first key of this sub-tile inside the partner's N keys
      float xmax = -INFINITY;
      uint64_t a1 = 0, a2 = 0;
#pragma unroll
      for (int c = 0; c < kSubN; c += 2) {
        if (c % FEAT == 0) {               // start of a key-grid row inside the sub-tile (FEAT >= 32: c == 0 only)
          const int key = key0 + c;
          const float t0 = fmaf(As, (float)(key % FEAT), fmaf(Bs, (float)(key / FEAT), base));
          a1 = pack2(bm - t0, bm - t0 - As);
          a2 = pack2(bm + t0, bm + t0 + As);
        }
        float a1x, a1y, a2x, a2y;
        unpack2(a1, a1x, a1y);
        unpack2(a2, a2x, a2y);
#ifdef CVD_WHATIF_NOBIAS
        const float v0 = negm, v1 = negm;
#else
        const float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
#endif
        float x0, x1;
        unpack2(fma2(pack2(__uint_as_float(sreg[c]), __uint_as_float(sreg[c + 1])), scale2, pack2(v0, v1)), x0, x1);
        if (N < kSubN && c >= N) x0 = x1 = -INFINITY;   // keys past the row (N = 16): not in the softmax
#ifndef CVD_WHATIF_NOMAX
        xmax = max3(xmax, x0, x1);
#else
        xmax = 0.f;
#endif
        sreg[c] = __float_as_uint(x0);
        sreg[c + 1] = __float_as_uint(x1);
        a1 = add2(a1, step1);
        a2 = add2(a2, step2);
      }
      if (DBG && stamps) stamps[3] = clock64() + (long long)(__float_as_uint(xmax) & 1u);

      // lazy maximum update (always on the very first sub-tile)
      const bool need = (t == 0) || (xmax > kRescaleThreshold);
      if (__any_sync(0xffffffffu, need)) {
        if (t > 0) {
          // O holds sub-tiles 0..t-1: PV(t-1) must have landed before O is touched.  (A thread is never
          // two phases behind pv_done: PV(t) cannot complete before this thread's own p_full arrival.)
          mbar_wait(smem_u32(&bars->pv_done), (t - 1) & 1);
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
        }
        if (need) {
          // First sub-tile: anchor the reference at the largest score this row can still reach through
          // the bias (the epipolar band may lie many key rows ahead), so that approaching the band does
          // not trigger a rescale every other sub-tile.  Any reference >= the running maximum is valid.
          const float shift = (t == 0) ? fmaxf(xmax, bias_ub) : xmax;
          m_used += shift;
#pragma unroll
          for (int c = 0; c < kSubN; ++c) sreg[c] = __float_as_uint(__uint_as_float(sreg[c]) - shift);
        }
      }
      uint32_t preg[kSubN / 2];
#pragma unroll
      for (int c = 0; c < kSubN; c += 2)
#ifdef CVD_WHATIF_NOEXP
        preg[c >> 1] = pack_bf16(__uint_as_float(sreg[c]) * 0.01f, __uint_as_float(sreg[c + 1]) * 0.01f);
#else
        preg[c >> 1] = pack_bf16(ex2(__uint_as_float(sreg[c])), ex2(__uint_as_float(sreg[c + 1])));
#endif
      if (DBG && stamps) stamps[4] = clock64() + (long long)(preg[0] & 1u);
      tmem_st_x16(tS, preg);               // P_t over the first 16 columns of the S buffer it came from
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full[t & 1]));
      if (DBG && stamps) stamps[5] = clock64();
    }

    // epilogue: O[:, 0:D] / O[:, D]  ->  out[r, n, h*D : (h+1)*D]
    mbar_wait(smem_u32(&bars->o_done), 0);   // committed once, after the last PV
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
  using C = Cfg<D, FEAT>;
  constexpr int N = FEAT * FEAT;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_map(&tmQ, q, (uint64_t)R * N, (uint64_t)H * D, kBlockM)) != CVD_OK) return rc;
  if ((rc = make_map(&tmK, k, (uint64_t)R_kv * N, (uint64_t)H * D, kLoadN)) != CVD_OK) return rc;
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
