// Opt-in A/B variants of the d = 40 forward kernel (included by epi_attn_sm100.cu inside namespace cvd::sm100, after the
// production kernel and the tensor-map helpers).  Both are exact (forced-variant parity tests) and both measured SLOWER
// than the production kernel at the bench shape (222 us): ping-pong 240 us (CVD_ATTN_PP=1), column split 254 us
// (CVD_ATTN_CS=1).  They are kept because their measurements are the evidence for what does NOT bound the kernel:
// hand-off gaps (ping-pong hides them: 13 % waiting, no gain) and warps per scheduler (column split doubles them: no
// gain).  profiles/r01_summary.md has the numbers.
#pragma once

// ------------------------------------------------------------------------------------------
// d = 40 "ping-pong" variant ("pp").  Measured on the kernel above (scripts/timeline_attn_v4.py, profiles/): its three
// chains per SM share the MUFU pipe fairly, so they finish their tiles together and then sit in their hand-offs
// (P -> issuer -> P V -> Q K^T -> S, 1000-1400 cycles) together; the four softmax warps of a chain also wait for the
// slowest of them at every hand-off.  Here every softmax warp owns TWO chains -- two 128-query tiles of the same
// (row, head) -- and alternates between them: while chain A's P V_j and Q K_{j+1}^T are in flight the warp computes
// chain B's tile j, so a warp only waits when a hand-off takes longer than a whole tile of arithmetic.
//   * CTA = (row, head, PAIR of query tiles), 160 threads: warps 0-3 softmax (thread = query row of both tiles), warp 4
//     TMA producer + MMA issuer.  Two CTAs per SM: 8 softmax warps that are (almost) always in their arithmetic.
//   * TMEM (256 columns): S_A [0,64) S_B [64,128) O_A [128,176) O_B [176,224); P_c is written in place over S_c (bf16,
//     32 columns) once S_c is in registers, so Q K_{j+1}^T of a chain follows its P V_j in the issue order.
//   * both chains read the same K / V^T tiles: one TMA load per tile and CTA serves 256 queries.
//   * softmax arithmetic as in the fused loop above (exponentials inside the score block, vote after them).
// ------------------------------------------------------------------------------------------
template <int D, int FEAT>
struct CfgP {
  static_assert(D <= 48 && D % 8 == 0, "pp variant: head_dim <= 48");
  static constexpr int kN = FEAT * FEAT;
  static constexpr int kTiles = kN / kTileN;
  static constexpr int kQTiles = kN / kBlockM;
  static_assert(kN % kTileN == 0 && kN % (2 * kBlockM) == 0, "pp variant: whole key tiles and an even number of query tiles");
  static constexpr int kKSteps = (D + 15) / 16;
  static constexpr int kDV = (D / 16 + 1) * 16;
  static constexpr int kQBytes = kBlockM * 128, kKBytes = kTileN * 128, kVBytes = kDV * 128;
  static constexpr int kColO = 2 * kTileN;                       // S_c at c*64, O_c at 128 + c*kDV
  static_assert(kColO + 2 * kDV <= 256, "TMEM budget");
  static constexpr int kStages = kTiles < 4 ? kTiles : 4;
  static constexpr int kSmemBytes = 2 * kQBytes + kStages * (kKBytes + kVBytes) + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct BarriersP {
  uint64_t q_full, s_full[2], p_full[2], o_done[2];
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint32_t tmem_base;
};

template <int D, int FEAT>
__global__ void __launch_bounds__(kThreads, 2)
epi_attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const float* __restrict__ rowp,
                       const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, int R, int H, int n_partners,
                       float step, float* __restrict__ lse, long long* __restrict__ prof) {
  using C = CfgP<D, FEAT>;
  constexpr int N = C::kN;
  constexpr int kPairs = C::kQTiles / 2;
  constexpr int kRowsPerTile = (kTileN + FEAT - 1) / FEAT;
  constexpr uint32_t kIdescQK = umma_idesc_bf16(kBlockM, kTileN);
  constexpr uint32_t kIdescPV = umma_idesc_bf16(kBlockM, C::kDV);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + 2 * C::kQBytes;
  const uint32_t sV = sK + C::kStages * C::kKBytes;
  BarriersP* bars = reinterpret_cast<BarriersP*>(smem_raw + (sV + C::kStages * C::kVBytes - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x % kPairs;
  const int h = (blockIdx.x / kPairs) % H;
  const int r = blockIdx.x / (kPairs * H);
  const int total = n_partners * C::kTiles;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int c = 0; c < 2; ++c) {
      mbar_init(smem_u32(&bars->s_full[c]), 1);
      mbar_init(smem_u32(&bars->p_full[c]), 128);
      mbar_init(smem_u32(&bars->o_done[c]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    __syncwarp();
    tmem_alloc<256>(smem_u32(&bars->tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    // ===================== TMA producer + MMA issuer =====================
    auto load_k = [&](int lt) {
      const int p = lt / C::kTiles, jt = lt % C::kTiles;
      const int rr = partner[p * R + r];
      const int st = lt % C::kStages;
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->k_full[st]), C::kKBytes);
        tma_load_3d(sK + st * C::kKBytes, &tmK, 0, h, rr * N + jt * kTileN, smem_u32(&bars->k_full[st]));
      }
      __syncwarp();
    };
    auto load_v = [&](int lv) {
      const int p = lv / C::kTiles, jt = lv % C::kTiles;
      const int rr = partner[p * R + r];
      const int st = lv % C::kStages;
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->v_full[st]), C::kVBytes);
        tma_load_2d(sV + st * C::kVBytes, &tmV, jt * kTileN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]));
      }
      __syncwarp();
    };
    // Q K_j^T of chain c (K stage already known to be full); the K stage is released behind chain 1's MMAs
    auto issue_qk = [&](int j, int c) {
      const int st = j % C::kStages;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ++ks)
          umma_ss(tmem + c * kTileN, umma_desc_k_sw128(sQ + c * C::kQBytes + ks * 32),
                  umma_desc_k_sw128(sK + st * C::kKBytes + ks * 32), kIdescQK, ks > 0);
        umma_commit(smem_u32(&bars->s_full[c]));
        if (c == 1) umma_commit(smem_u32(&bars->k_empty[st]));
      }
      __syncwarp();
    };
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bars->q_full), 2 * C::kQBytes);
      tma_load_3d(sQ, &tmQ, 0, h, r * N + (2 * qp) * kBlockM, smem_u32(&bars->q_full));
      tma_load_3d(sQ + C::kQBytes, &tmQ, 0, h, r * N + (2 * qp + 1) * kBlockM, smem_u32(&bars->q_full));
    }
    __syncwarp();
    for (int lt = 0; lt < C::kStages && lt < total; ++lt) load_k(lt);
    for (int lv = 0; lv < C::kStages && lv < total; ++lv) load_v(lv);
    mbar_wait(smem_u32(&bars->q_full), 0);
    mbar_wait(smem_u32(&bars->k_full[0]), 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(0, 1);
    long long t_ref = 0, t_pw = 0, t_pv = 0, t_qk = 0, t0 = 0;
    for (int j = 0; j < total; ++j) {
      const int st = j % C::kStages;
      if (prof) t0 = clock64();
      // refills: the consumers of tile j-1 were issued a whole iteration ago
      if (j >= 1 && j - 1 + C::kStages < total) {
        mbar_wait(smem_u32(&bars->k_empty[(j - 1) % C::kStages]), ((j - 1) / C::kStages) & 1);
        load_k(j - 1 + C::kStages);
        mbar_wait(smem_u32(&bars->v_empty[(j - 1) % C::kStages]), ((j - 1) / C::kStages) & 1);
        load_v(j - 1 + C::kStages);
      }
      mbar_wait(smem_u32(&bars->v_full[st]), (j / C::kStages) & 1);
      if (j + 1 < total) mbar_wait(smem_u32(&bars->k_full[(j + 1) % C::kStages]), ((j + 1) / C::kStages) & 1);
      if (prof) { const long long t1 = clock64(); t_ref += t1 - t0; t0 = t1; }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mbar_wait(smem_u32(&bars->p_full[c]), j & 1);
        tc_fence_after();
        if (prof) { const long long t1 = clock64(); t_pw += t1 - t0; t0 = t1; }
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kTileN / 16; ++ks)
            umma_ts(tmem + C::kColO + c * C::kDV, tmem + c * kTileN + ks * 8, umma_desc_k_sw128(sV + st * C::kVBytes + ks * 32),
                    kIdescPV, (j > 0) || (ks > 0));
          if (j == total - 1) umma_commit(smem_u32(&bars->o_done[c]));
          if (c == 1) umma_commit(smem_u32(&bars->v_empty[st]));
        }
        __syncwarp();
        if (prof) { const long long t1 = clock64(); t_pv += t1 - t0; t0 = t1; }
        if (j + 1 < total) issue_qk(j + 1, c);     // tcgen05.mma executes in issue order: S_c is rewritten after P V_j read P_c
        if (prof) { const long long t1 = clock64(); t_qk += t1 - t0; t0 = t1; }
      }
    }
    if (prof && lane == 0) {
      long long* o = prof + (size_t)blockIdx.x * 16;
      o[8] = t_ref; o[9] = t_pw; o[10] = t_pv; o[11] = t_qk;
    }
  } else {
    // =============================== softmax warps ==============================
    const int tid = threadIdx.x;                       // query row inside both tiles == TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float off = (step - 1.0f) * 0.5f;
    const float scale_log2 = kLog2e * rsqrtf((float)D);
    const uint64_t scale2 = pack2(scale_log2, scale_log2);
    // per chain: bias geometry of the query pixel (base-2 units) and the running reference maximum
    float xq[2], yq[2], As[2] = {0.f, 0.f}, Bs[2] = {0.f, 0.f}, base[2] = {0.f, 0.f}, m_used[2] = {0.f, 0.f},
          bias_ub[2] = {0.f, 0.f};
    float bandp = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int n = (2 * qp + c) * kBlockM + tid;
      xq[c] = pixel_centre(n % FEAT, step);
      yq[c] = pixel_centre(n / FEAT, step);
    }

    long long t_wait = 0, t_loop = prof ? clock64() : 0;
    for (int j = 0; j < total; ++j) {
      const int p = j / C::kTiles, jj = j % C::kTiles;
      if (jj == 0 && rowp != nullptr) {
        const float* P = rowp + (size_t)(p * R + r) * CVD_ROWP_STRIDE;
        bandp = P[9] * P[10] * kLog2e;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float a = fmaf(P[1], yq[c], P[0] * xq[c]) + P[2];
          const float b = fmaf(P[4], yq[c], P[3] * xq[c]) + P[5];
          const float cc = fmaf(P[7], yq[c], P[6] * xq[c]) + P[8];
          const float k1 = P[10] * kLog2e / (sqrtf(a * a + b * b) + 1e-6f);
          const float A = a * k1, B = b * k1;
          As[c] = A * step;
          Bs[c] = B * step;
          base[c] = fmaf(A, off, fmaf(B, off, cc * k1));
          if (j == 0) {     // upper bound of the bias over the key grid (|t| is convex: corners)
            const float e = (float)(FEAT - 1);
            const float t00 = base[c], t10 = fmaf(As[c], e, base[c]), t01 = fmaf(Bs[c], e, base[c]), t11 = fmaf(As[c], e, t01);
            const float lo = fminf(fminf(t00, t10), fminf(t01, t11)), hi = fmaxf(fmaxf(t00, t10), fmaxf(t01, t11));
            const float tmin = (lo <= 0.f && hi >= 0.f) ? 0.f : fminf(fabsf(lo), fabsf(hi));
            bias_ub[c] = fminf(bandp - tmin, 0.f);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t tS = tmem + lane_off + c * kTileN, tO = tmem + lane_off + C::kColO + c * C::kDV;
        const long long tw0 = prof ? clock64() : 0;
        mbar_wait(smem_u32(&bars->s_full[c]), j & 1);     // implies P V_{j-1} of this chain has completed
        tc_fence_after();
        if (prof) t_wait += clock64() - tw0;
        constexpr int kQ = 16;
        uint32_t preg[kTileN / 2];
        uint32_t s0[kHalfN], s1[kHalfN];
        tmem_ld_x32(tS, s0);
        tmem_ld_wait();
        const uint64_t step1 = pack2(-2.f * As[c], -2.f * As[c]), step2 = pack2(2.f * As[c], 2.f * As[c]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t* cur = (q < 2 ? s0 : s1) + (q & 1) * kQ;
          if (q == 0) tmem_ld_x32(tS + kHalfN, s1);          // lands under this block
          const float bm = bandp - m_used[c], negm = -m_used[c];
          float xmax = -INFINITY;
          uint64_t a1 = 0, a2 = 0;
#pragma unroll
          for (int i = 0; i < kQ; i += 2) {
            const int cc = q * kQ + i;           // key index inside the 64-key tile
            if (cc % FEAT == 0 || i == 0) {      // start of a key-grid row (or of this quarter)
              const float rb = fmaf(Bs[c], (float)(jj * kRowsPerTile + cc / FEAT), base[c]);
              const float t0 = fmaf(As[c], (float)(cc % FEAT), rb);
              a1 = pack2(bm - t0, bm - t0 - As[c]);
              a2 = pack2(bm + t0, bm + t0 + As[c]);
            }
            float a1x, a1y, a2x, a2y;
            unpack2(a1, a1x, a1y);
            unpack2(a2, a2x, a2y);
            const float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
            float x0, x1;
            unpack2(fma2(pack2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), scale2, pack2(v0, v1)), x0, x1);
            xmax = max3(xmax, x0, x1);
            preg[q * (kQ / 2) + (i >> 1)] = pack_bf16(ex2(x0), ex2(x1));
            cur[i] = __float_as_uint(x0);
            cur[i + 1] = __float_as_uint(x1);
            a1 = add2(a1, step1);
            a2 = add2(a2, step2);
          }
          if (q == 0) tmem_ld_wait_dep32(s1);
          // lazy maximum update (always on the very first quarter)
          const bool first = (j == 0 && q == 0);
          const bool need = first || (xmax > kRescaleThreshold);
          if (__any_sync(0xffffffffu, need)) {
            if (!first) {
              const float f = need ? ex2(-xmax) : 1.0f;
              if (j > 0) {          // O_c holds tiles 0..j-1 (complete: see s_full above)
#pragma unroll
                for (int c0 = 0; c0 < C::kDV; c0 += 16) {
                  uint32_t o[16];
                  tmem_ld_x16(tO + c0, o);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                  tmem_st_x16(tO + c0, o);
                }
              }
#pragma unroll
              for (int i = 0; i < q * (kQ / 2); ++i)
                preg[i] = pack_bf16(__uint_as_float(preg[i] << 16) * f, __uint_as_float(preg[i] & 0xffff0000u) * f);
            }
            const float shift = need ? (first ? fmaxf(xmax, bias_ub[c]) : xmax) : 0.f;
            m_used[c] += shift;
#pragma unroll
            for (int i = 0; i < kQ; i += 2)
              preg[q * (kQ / 2) + (i >> 1)] =
                  pack_bf16(ex2(__uint_as_float(cur[i]) - shift), ex2(__uint_as_float(cur[i + 1]) - shift));
          }
        }
        tmem_st_x32(tS, preg);          // P_c in place over S_c (all 64 scores are in registers)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[c]));
      }
    }

    if (prof && lane == 0) {
      long long* o = prof + (size_t)blockIdx.x * 16;
      o[warp * 2] = t_wait; o[warp * 2 + 1] = clock64() - t_loop;
    }
    // epilogue: O_c[:, 0:D] / O_c[:, D]  ->  out[r, n, h*D : (h+1)*D]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t tO = tmem + lane_off + C::kColO + c * C::kDV;
      const int n = (2 * qp + c) * kBlockM + tid;
      mbar_wait(smem_u32(&bars->o_done[c]), 0);
      tc_fence_after();
      uint32_t lraw;
      tmem_ld_x1(tO + D, lraw);
      tmem_ld_wait();
      const float inv = 1.0f / __uint_as_float(lraw);
      if (lse != nullptr) lse[((size_t)r * H + h) * N + n] = m_used[c] + log2f(__uint_as_float(lraw));
      __nv_bfloat16* orow = out + ((size_t)r * N + n) * (size_t)(H * D) + h * D;
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 8) {
        uint32_t o[8];
        tmem_ld_x8(tO + c0, o);
        tmem_ld_wait();
        uint4 pk;
        pk.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        pk.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        pk.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        pk.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        *reinterpret_cast<uint4*>(orow + c0) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}


// ------------------------------------------------------------------------------------------
// d = 40 "column split" variant ("cs").  The softmax loop is bound by dependent-issue latency with 3-4 warps per
// scheduler (profiles/: `wait` 1.6 stall cycles per issued instruction, XU 58 %), and TMEM caps the chains at three per
// SM.  Here the 128 x 64 score tile of a chain is shared by TWO warpgroups: warp w and warp w + 4 own the same 32 TMEM
// lanes (query rows) and 32 of the 64 key columns each, so three CTAs put 24 softmax warps on an SM (6 per scheduler)
// with the same TMEM footprint (S 64 + O 48 + P 32 columns per CTA).
//   * the two threads of a query row must scale P with the SAME reference maximum (one O accumulator): after its 32
//     exponentials each thread publishes "my largest x exceeds the reference by more than 2^16" (or the anchor request
//     of the very first tile) in shared memory, the warp pair meets at a 64-thread named barrier, and if either asked,
//     BOTH move the reference by the same amount and simply redo the tile (S is still in TMEM: s_free is signalled
//     after the decision).  Rare after the first tile; overflow-proof, because nothing computed before a redo survives.
//   * warps 0-7 softmax (lane quarter = warp & 3, column half = warp >> 2), warp 8 TMA producer + MMA issuer.
// ------------------------------------------------------------------------------------------
constexpr int kThreadsC = 288;

struct BarriersC {
  uint64_t q_full, s_full, s_free, p_full, pv_done, o_done;
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint32_t tmem_base, tmem_p;
  float xch[2][2][kBlockM];     // [slot][column half][query row]: reference-move requests of the warp pairs
};

template <int D, int FEAT>
struct CfgC {
  using B = Cfg<D, FEAT>;
  static_assert(B::kColO + B::kDV <= 128 && B::kKBlocks == 1, "cs variant: d <= 48 (S + O in 128 columns, P in its own 32)");
  static_assert(B::kN % kTileN == 0, "cs variant: whole key tiles");
  // three CTAs per SM (128 + 32 TMEM columns each): a fourth must not fit
  static constexpr int kSmemBytes = (B::kSmemData < 58 * 1024 ? 58 * 1024 : B::kSmemData) + 1024 + 256 + 2048;
};

template <int D, int FEAT>
__global__ void __launch_bounds__(kThreadsC, 3)
epi_attn_fwd_cs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const float* __restrict__ rowp,
                       const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, int R, int H, int n_partners,
                       float step, float* __restrict__ lse) {
  using C = Cfg<D, FEAT>;
  constexpr int N = C::kN;
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  constexpr int kRowsPerTile = (kTileN + FEAT - 1) / FEAT;
  constexpr uint32_t kIdescQK = umma_idesc_bf16(kBlockM, kTileN);
  constexpr uint32_t kIdescPV = umma_idesc_bf16(kBlockM, C::kDV);

  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + C::kQBytes;
  const uint32_t sV = sK + C::kKStages * C::kKBytes;
  BarriersC* bars = reinterpret_cast<BarriersC*>(smem_raw + (sV + C::kVStages * C::kVBytes - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % kQTiles;
  const int h = (blockIdx.x / kQTiles) % H;
  const int r = blockIdx.x / (kQTiles * H);
  const int total = n_partners * C::kTiles;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(smem_u32(&bars->q_full), 1);
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_free), 256);
    mbar_init(smem_u32(&bars->p_full), 256);
    mbar_init(smem_u32(&bars->pv_done), 1);
    mbar_init(smem_u32(&bars->o_done), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    __syncwarp();
    tmem_alloc_cols<C::kTmemCols>(smem_u32(&bars->tmem_base));
    tmem_alloc_cols<32>(smem_u32(&bars->tmem_p));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tmemP = bars->tmem_p;

  if (warp == 8) {
    // ===================== TMA producer + MMA issuer (as in epi_attn_fwd_kernel) =====================
    auto load_k = [&](int lt) {
      const int p = lt / C::kTiles, jt = lt % C::kTiles;
      const int rr = partner[p * R + r];
      const int st = lt % C::kKStages;
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->k_full[st]), C::kKBytes);
        tma_load_3d(sK + st * C::kKBytes, &tmK, 0, h, rr * N + jt * kTileN, smem_u32(&bars->k_full[st]));
      }
      __syncwarp();
    };
    auto load_v = [&](int lv) {
      const int p = lv / C::kTiles, jt = lv % C::kTiles;
      const int rr = partner[p * R + r];
      const int st = lv % C::kVStages;
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->v_full[st]), C::kVBytes);
        tma_load_2d(sV + st * C::kVBytes, &tmV, jt * kTileN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]));
      }
      __syncwarp();
    };
    auto issue_qk = [&](int j) {
      const int st = j % C::kKStages;
      mbar_wait(smem_u32(&bars->k_full[st]), (j / C::kKStages) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ++ks)
          umma_ss(tmem + C::kColS, umma_desc_k_sw128(sQ + ks * 32), umma_desc_k_sw128(sK + st * C::kKBytes + ks * 32),
                  kIdescQK, ks > 0);
        umma_commit(smem_u32(&bars->s_full));
        umma_commit(smem_u32(&bars->k_empty[st]));
      }
      __syncwarp();
    };
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::kQBytes);
      tma_load_3d(sQ, &tmQ, 0, h, r * N + qt * kBlockM, smem_u32(&bars->q_full));
    }
    __syncwarp();
    for (int lt = 0; lt < C::kKStages && lt < total; ++lt) load_k(lt);
    for (int lv = 0; lv < C::kVStages && lv < total; ++lv) load_v(lv);
    mbar_wait(smem_u32(&bars->q_full), 0);
    issue_qk(0);
    for (int j = 0; j < total; ++j) {
      const int st = j % C::kVStages;
      if (j + C::kKStages < total) {
        mbar_wait(smem_u32(&bars->k_empty[j % C::kKStages]), (j / C::kKStages) & 1);
        load_k(j + C::kKStages);
      }
      if (j + 1 < total) {
        mbar_wait(smem_u32(&bars->s_free), j & 1);
        issue_qk(j + 1);
      }
      if (j >= 1 && j - 1 + C::kVStages < total) {
        mbar_wait(smem_u32(&bars->v_empty[(j - 1) % C::kVStages]), ((j - 1) / C::kVStages) & 1);
        load_v(j - 1 + C::kVStages);
      }
      mbar_wait(smem_u32(&bars->v_full[st]), (j / C::kVStages) & 1);
      mbar_wait(smem_u32(&bars->p_full), j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kTileN / 16; ++ks)
          umma_ts(tmem + C::kColO, tmemP + ks * 8, umma_desc_k_sw128(sV + st * C::kVBytes + ks * 32), kIdescPV,
                  (j > 0) || (ks > 0));
        umma_commit(smem_u32(&bars->pv_done));
        if (j == total - 1) umma_commit(smem_u32(&bars->o_done));
        umma_commit(smem_u32(&bars->v_empty[st]));
      }
      __syncwarp();
    }
  } else {
    // =============================== softmax warps ==============================
    const int half = warp >> 2;                          // which 32 of the 64 key columns
    const int row = (warp & 3) * 32 + lane;              // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS = tmem + lane_off + C::kColS + half * 32, tO = tmem + lane_off + C::kColO;
    const uint32_t tP = tmemP + lane_off + half * 16;
    const int n = qt * kBlockM + row;
    const float xq = pixel_centre(n % FEAT, step), yq = pixel_centre(n / FEAT, step);
    const float off = (step - 1.0f) * 0.5f;
    const float scale_log2 = kLog2e * rsqrtf((float)D);
    const uint64_t scale2 = pack2(scale_log2, scale_log2);
    float As = 0.f, Bs = 0.f, base = 0.f, bandp = 0.f, m_used = 0.f, bias_ub = 0.f;
    bool anchored = false;
    int xslot = 0;

    for (int j = 0; j < total; ++j) {
      const int p = j / C::kTiles, jj = j % C::kTiles;
      if (jj == 0 && rowp != nullptr) {
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
        if (j == 0) {
          const float e = (float)(FEAT - 1);
          const float t00 = base, t10 = fmaf(As, e, base), t01 = fmaf(Bs, e, base), t11 = fmaf(As, e, t01);
          const float lo = fminf(fminf(t00, t10), fminf(t01, t11)), hi = fmaxf(fmaxf(t00, t10), fmaxf(t01, t11));
          const float tmin = (lo <= 0.f && hi >= 0.f) ? 0.f : fminf(fabsf(lo), fabsf(hi));
          bias_ub = fminf(bandp - tmin, 0.f);
        }
      }
      mbar_wait(smem_u32(&bars->s_full), j & 1);
      tc_fence_after();

      uint32_t preg[16];
      const uint64_t step1 = pack2(-2.f * As, -2.f * As), step2 = pack2(2.f * As, 2.f * As);
      bool redo;
      do {
        const float bm = bandp - m_used, negm = -m_used;
        float xmax = -INFINITY;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t sr[16];
          tmem_ld_x16(tS + q * 16, sr);
          tmem_ld_wait();
          uint64_t a1 = 0, a2 = 0;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const int cc = half * 32 + q * 16 + i;      // key index inside the 64-key tile (compile-time up to `half`)
            if (i == 0 || (FEAT < 16 && (q * 16 + i) % FEAT == 0)) {      // start of this quarter (or of a key-grid row)
              const float rb = fmaf(Bs, (float)(jj * kRowsPerTile + cc / FEAT), base);
              const float t0 = fmaf(As, (float)(cc % FEAT), rb);
              a1 = pack2(bm - t0, bm - t0 - As);
              a2 = pack2(bm + t0, bm + t0 + As);
            }
            float a1x, a1y, a2x, a2y;
            unpack2(a1, a1x, a1y);
            unpack2(a2, a2x, a2y);
            const float v0 = min3(a1x, a2x, negm), v1 = min3(a1y, a2y, negm);
            float x0, x1;
            unpack2(fma2(pack2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), scale2, pack2(v0, v1)), x0, x1);
            xmax = max3(xmax, x0, x1);
            preg[q * 8 + (i >> 1)] = pack_bf16(ex2(x0), ex2(x1));
            a1 = add2(a1, step1);
            a2 = add2(a2, step2);
          }
        }
        // the two threads of a query row agree on the reference: publish my request, meet my partner warp, read its
        const bool need = !anchored || (xmax > kRescaleThreshold);
        bars->xch[xslot][half][row] = need ? xmax : -INFINITY;
        switch (warp & 3) {      // immediate barrier ids: a register id would reserve all 16 named barriers per CTA
          case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
          case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
          case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
          default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
        }
        const float other = bars->xch[xslot][half ^ 1][row];
        xslot ^= 1;
        const float m = fmaxf(need ? xmax : -INFINITY, other);
        redo = __any_sync(0xffffffffu, m > -INFINITY);       // identical in both warps of the pair (m is per row)
        if (redo) {
          const float shift = (m > -INFINITY) ? (anchored ? m : fmaxf(m, bias_ub)) : 0.f;
          if (j > 0 && half == 0) {       // O holds tiles 0..j-1
            mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
            tc_fence_after();
            const float f = ex2(-shift);
#pragma unroll
            for (int c0 = 0; c0 < C::kDV; c0 += 16) {
              uint32_t o[16];
              tmem_ld_x16(tO + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x16(tO + c0, o);
            }
            tmem_st_wait();
          }
          m_used += shift;
          anchored = true;
        }
      } while (redo);
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_free));      // S_j will not be read again: Q K_{j+1}^T may overwrite it
      if (j > 0) {   // the P columns are free once PV(j-1) has read them
        mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_x16(tP, preg);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full));
    }

    // epilogue: the two threads of a row share the read-out (chunks of 8 output columns, alternating)
    mbar_wait(smem_u32(&bars->o_done), 0);
    tc_fence_after();
    uint32_t lraw;
    tmem_ld_x1(tO + D, lraw);
    tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(lraw);
    if (lse != nullptr && half == 0) lse[((size_t)r * H + h) * N + n] = m_used + log2f(__uint_as_float(lraw));
    __nv_bfloat16* orow = out + ((size_t)r * N + n) * (size_t)(H * D) + h * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 8) {
      if (((c0 >> 3) & 1) != half) continue;
      uint32_t o[8];
      tmem_ld_x8(tO + c0, o);
      tmem_ld_wait();
      uint4 pk;
      pk.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      pk.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      pk.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      pk.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      *reinterpret_cast<uint4*>(orow + c0) = pk;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem);
    tmem_dealloc<32>(bars->tmem_p);
  }
}

template <int D, int FEAT>
static int launch_cs(const void* q, const void* k, const void* vt, const float* rowp, const int* partner, void* out, int R,
                     int R_kv, int H, int n_partners, float S, float* lse, cudaStream_t st) {
  using C = Cfg<D, FEAT>;
  constexpr int N = FEAT * FEAT;
  static_assert(N % kBlockM == 0, "cs variant: whole query tiles");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_map_heads(&tmQ, q, (uint64_t)R * N, (uint64_t)H, (uint64_t)D, kBlockM)) != CVD_OK) return rc;
  if ((rc = make_map_heads(&tmK, k, (uint64_t)R_kv * N, (uint64_t)H, (uint64_t)D, kTileN)) != CVD_OK) return rc;
  if ((rc = make_map(&tmV, vt, (uint64_t)R_kv * H * C::kDV, (uint64_t)N, C::kDV)) != CVD_OK) return rc;
  static bool configured = false;
  if (!configured) {
    CVD_CUDA(cudaFuncSetAttribute(epi_attn_fwd_cs_kernel<D, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  CfgC<D, FEAT>::kSmemBytes));
    configured = true;
  }
  const long long grid = (long long)R * H * (N / kBlockM);
  CVD_REQUIRE(grid > 0 && grid < (1ll << 31), "cvd_epi_attn_fwd: grid of %lld CTAs", grid);
  epi_attn_fwd_cs_kernel<D, FEAT><<<(unsigned)grid, kThreadsC, CfgC<D, FEAT>::kSmemBytes, st>>>(
      tmQ, tmK, tmV, rowp, partner, (__nv_bfloat16*)out, R, H, n_partners, S / (float)FEAT, lse);
  CVD_LAUNCH_CHECK("epi_attn_fwd_cs_kernel");
  return CVD_OK;
}

static long long* g_pp_prof = nullptr;
static long long g_pp_prof_ctas = 0, g_pp_prof_used = 0;

template <int D, int FEAT>
static int launch_pp(const void* q, const void* k, const void* vt, const float* rowp, const int* partner, void* out, int R,
                     int R_kv, int H, int n_partners, float S, float* lse, cudaStream_t st) {
  using C = CfgP<D, FEAT>;
  constexpr int N = FEAT * FEAT;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_map_heads(&tmQ, q, (uint64_t)R * N, (uint64_t)H, (uint64_t)D, kBlockM)) != CVD_OK) return rc;
  if ((rc = make_map_heads(&tmK, k, (uint64_t)R_kv * N, (uint64_t)H, (uint64_t)D, kTileN)) != CVD_OK) return rc;
  if ((rc = make_map(&tmV, vt, (uint64_t)R_kv * H * C::kDV, (uint64_t)N, C::kDV)) != CVD_OK) return rc;
  static bool configured = false;
  if (!configured) {
    CVD_CUDA(cudaFuncSetAttribute(epi_attn_fwd_pp_kernel<D, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const long long grid = (long long)R * H * (C::kQTiles / 2);
  CVD_REQUIRE(grid > 0 && grid < (1ll << 31), "cvd_epi_attn_fwd: grid of %lld CTAs", grid);
  // CVD_ATTN_PP_PROF=1 (profiling only): per-CTA cycle counters of the waits and issue phases, printed by
  // cvd_epi_attn_pp_profile_dump()
  long long* prof = nullptr;
  if (getenv("CVD_ATTN_PP_PROF") != nullptr) {
    if (g_pp_prof == nullptr || g_pp_prof_ctas < grid) {
      if (g_pp_prof) cudaFree(g_pp_prof);
      CVD_CUDA(cudaMalloc(&g_pp_prof, (size_t)grid * 16 * sizeof(long long)));
      g_pp_prof_ctas = grid;
    }
    CVD_CUDA(cudaMemsetAsync(g_pp_prof, 0, (size_t)grid * 16 * sizeof(long long), st));
    prof = g_pp_prof;
    g_pp_prof_used = grid;
  }
  epi_attn_fwd_pp_kernel<D, FEAT><<<(unsigned)grid, kThreads, C::kSmemBytes, st>>>(tmQ, tmK, tmV, rowp, partner,
                                                                                (__nv_bfloat16*)out, R, H, n_partners,
                                                                                S / (float)FEAT, lse, prof);
  CVD_LAUNCH_CHECK("epi_attn_fwd_pp_kernel");
  return CVD_OK;
}

