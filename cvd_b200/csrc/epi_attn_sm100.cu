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
//  * S = Q K^T and O += P V run on the 5th-gen tensor cores (tcgen05.mma, cta_group::1, M=128,
//    64 keys per round): Q/K/V^T tiles arrive by TMA (128B swizzle) into shared memory, S, P and
//    O live in TMEM, P is the TMEM A operand of the second MMA (never through shared memory),
//    and the softmax denominator is a column of O (V^T carries a row of ones).
//  * Measured on B200 (profiles/r01_summary.md): at d=40 the kernel sits between the hand-off latencies (every mbarrier /
//    TMEM round trip costs 100-350 cycles, a tcgen05.mma / commit 50-100 cycles of the issuing warp) and the softmax
//    instruction stream (11-12 scheduler cycles per warp-column in the kernel, MUFU floor 8).  Hence: few, 64-key MMA
//    rounds, and as many chains per SM as TMEM and registers allow: d=40 runs FOUR CTAs per SM (S + O in one 128-column
//    allocation, P in place over S, 96 registers; see Cfg::kFour), d=80 two (P in its own columns, so Q K^T of tile j+1 is
//    issued as soon as S_j has been read and overlaps the softmax of tile j), d=160 one or two.
//  * Warp roles per CTA (160 threads): warps 0-3 softmax (one TMEM lane quarter each), warp 4 TMA producer + MMA issuer +
//    TMEM owner (one elected lane issues).  Five warps, not six: the per-scheduler warp count sets the register budget
//    (20 warps per SM: 96 registers; 15: 128; the six-warp layout allowed 80 / 96 and the fused softmax loop spilled).
//  * Softmax loop: a tile is walked in four 16-key quarters; in each, every score goes bias -> FFMA2 -> MUFU.EX2 -> F2FP
//    in one straight-line block (ptxas interleaves the MUFU stream with the FADD2 / FMNMX3 stream only through such true
//    dependencies), and the reference-maximum vote follows the exponentials.
//  * Online softmax with a lazily updated reference maximum anchored at the largest score the
//    row can still reach through the bias: scores are evaluated directly against the reference
//    (the subtraction is folded into the bias terms) and O is rescaled only when a score
//    exceeds it by more than 2^16 (rare).
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cvd {
// csrc/epi_attn_small.cu: warp-level kernel for the N <= 64 grids; -1000 = not a small-grid configuration
int epi_attn_small_dispatch(const void* q, const KvSrc& kv, const float* rowp, const int* partner, void* out, float* lse, int R,
                            int heads, int head_dim, int n_partners, int feat, float S, cudaStream_t st);
namespace sm100 {

using namespace ptx;

constexpr int kBlockM = 128;   // queries per CTA (= TMEM lanes)
constexpr int kTileN = 64;     // keys per TMA tile / MMA round
constexpr int kHalfN = 32;     // scores a softmax thread holds in registers at a time
constexpr int kThreads = 160;   // warps 0-3 softmax | warp 4 TMA producer + MMA issuer + TMEM owner
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 16.0f;  // lazy rescale: only when a score exceeds the reference by > 2^16

// CL = CTAs per cluster.  CL > 1 ("multicast"): the CL query tiles of one (row, head) form a thread-block cluster and
// share every K / V^T tile -- ONE TMA load per tile, issued by CTA (tile % CL), is delivered to the same shared-memory
// stage of all CL CTAs (cp.async.bulk.tensor ... .multicast::cluster): K / V^T leave L2 once per (row, head) instead of
// once per query tile.  A stage is refilled once ALL CL CTAs have consumed it: their MMA commits arrive (multicast) on
// the stage's "empty" barrier of every CTA, which therefore counts CL arrivals.  Built for K / V^T in a PEER GPU's HBM
// (views sharded across GPUs: every byte would cross NVLink once, under the softmax) -- but multicast TMA with a source
// in peer memory never completes on this platform (see cluster_requested() below), so it is a LOCAL-memory variant.
template <int D, int FEAT, int CL = 1>
struct Cfg {
  static_assert(D % 8 == 0 && (D % 16 == 0 || D % 16 == 8), "head_dim must be a multiple of 8");
  static_assert(CL >= 1 && CL <= 8 && (CL & (CL - 1)) == 0, "cluster size: 1, 2, 4 or 8");
  static constexpr int kN = FEAT * FEAT;
  static constexpr int kTiles = (kN + kTileN - 1) / kTileN;      // per partner
  static constexpr int kKBlocks = (D + 63) / 64;          // 128-byte swizzle atoms along the head dim
  static constexpr int kKSteps = (D + 15) / 16;           // UMMA K steps of Q K^T
  static constexpr int kDV = (D / 16 + 1) * 16;           // rows of a V^T tile: d values, a row of ones, zero pad
  static constexpr int kQBytes = kKBlocks * kBlockM * 128;
  static constexpr int kKBytes = kKBlocks * kTileN * 128;
  static constexpr int kVBytes = kDV * 128;
  // TMEM: S [0,64) | O [64,64+DV) in the main allocation.  P (32 columns of packed bf16):
  //   d=40 : aliased onto S inside one 128-column allocation -> 4 CTAs per SM (see kFour below)
  //   d=80 : behind O inside a 256-column allocation
  //   d=160: aliased onto S (then Q K^T of tile j+1 has to follow P_j V_j)
  static constexpr int kColS = 0, kColO = kTileN;
  // d <= 48 ("kFour"): S 64 + O 48 columns fit ONE 128-column allocation if P is written in place over S, so FOUR CTAs
  // (16 softmax warps) share an SM at 96 registers per thread -- the fused softmax loop then takes its scores from TMEM
  // 16 at a time.  Measured at the bench shape: 204 us against 222 us for three CTAs with P in its own 32 columns
  // (128 registers, Q K^T of tile j+1 overlapping the softmax of tile j); 4096 CTAs are also 6.9 waves of 592 slots
  // instead of 9.2 of 444.  make EXTRA=-DCVD_D40_THREE_CTAS rebuilds the three-CTA configuration for A/B runs.
#ifdef CVD_D40_THREE_CTAS
  static constexpr bool kFour = false;
#else
  static constexpr bool kFour = (kColO + kDV) <= 128;
#endif
  static constexpr bool kAliasP = kFour || (kColO + kDV + kTileN / 2) > 256;
  static constexpr bool kSplitP = !kAliasP && (kColO + kDV) <= 128;
  static constexpr int kTmemCols = (kSplitP || kFour) ? 128 : 256;
  static constexpr int kColP = kAliasP ? kColS : (kSplitP ? 0 : kColO + kDV);   // kSplitP: relative to the 2nd allocation
  static_assert(kColO + kDV <= 256, "TMEM budget");
  // K/V ring depth (never deeper than the tile count).  Multicast build: tiles may come over NVLink (~2 us of latency
  // against a tile period of ~1.3 us per CTA at four CTAs per SM), so K gets a third stage: K_{j+3} is requested when
  // every CTA of the cluster has finished Q K_j^T and is needed three tile periods later; V_{j+1} is requested when
  // P V_{j-1} has completed everywhere and is needed two periods later.
  static constexpr int kWant = kFour ? 2 : (D <= 40 ? 3 : (D <= 80 ? 2 : (kTiles > 1 ? 2 : 1)));
  static constexpr int kWantK = (CL > 1 && kWant < 3) ? 3 : kWant;
  static constexpr int kKStages = kWantK < kTiles ? kWantK : kTiles;
  static constexpr int kVStages = kWant < kTiles ? kWant : kTiles;
  static constexpr int kSmemData = kQBytes + kKStages * kKBytes + kVStages * kVBytes;
  // three-CTA build: d=40 must NOT fit a 4th CTA (4 x 128 TMEM columns would starve the 32-column allocations);
  // four-CTA build: must not fit a 5th (no TMEM left for it)
  static constexpr int kSmemFloor = kSplitP ? 58 * 1024 : (kFour ? 47 * 1024 : 0);
  static constexpr int kSmemBytes = (kSmemData < kSmemFloor ? kSmemFloor : kSmemData) + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kMinBlocks = kFour ? 4 : (kSplitP ? 3 : (kSmemBytes <= 110 * 1024 ? 2 : 1));
};

struct Barriers {
  uint64_t q_full, s_full, s_free, p_full, pv_done, o_done;
  uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
  uint32_t tmem_base, tmem_p;
};

// K / V^T may be spread over several buffers of equal row count (one per source GPU when the views are sharded: the
// symmetric-memory copies of the peers, addressed through their NVLink mappings); partner rows are numbered globally,
// row p lives in buffer p / rows_per_src at local row p % rows_per_src.  A single local buffer is source 0.
constexpr int kMaxSrc = kMaxKvSrc;
struct KvMaps {
  CUtensorMap k[kMaxSrc];
  CUtensorMap v[kMaxSrc];
};

template <int D, int FEAT, bool DBG, bool PIPE, int CL>
__global__ void __launch_bounds__(kThreads, Cfg<D, FEAT, CL>::kMinBlocks)
epi_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ KvMaps kv,
                    const float* __restrict__ rowp, const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, int R,
                    int H, int n_partners, int rows_per_src, float step, float* __restrict__ dbg, float* __restrict__ lse) {
  using C = Cfg<D, FEAT, CL>;
  constexpr uint16_t kMcMask = (uint16_t)((1u << CL) - 1u);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr int N = C::kN;
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  constexpr int kRowsPerTile = (kTileN + FEAT - 1) / FEAT;   // key-grid rows touched by one key tile
  constexpr uint32_t kIdescQK = umma_idesc_bf16(kBlockM, kTileN);
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
  const int total = n_partners * C::kTiles;
  pdl_trigger();      // the next kernel of the stream may fill the slots this grid's tail leaves idle

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&kv.k[0]);
    prefetch_tmap(&kv.v[0]);
    mbar_init(smem_u32(&bars->q_full), 1);
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_free), 128);
    mbar_init(smem_u32(&bars->p_full), 128);
    mbar_init(smem_u32(&bars->pv_done), 1);
    mbar_init(smem_u32(&bars->o_done), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), CL);      // one (multicast) commit arrival per CTA of the cluster
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), CL);
    }
    fence_mbar_init();
  }
  if (warp == 4) {
    __syncwarp();
    tmem_alloc_cols<C::kTmemCols>(smem_u32(&bars->tmem_base));
    if (C::kSplitP) tmem_alloc_cols<32>(smem_u32(&bars->tmem_p));
    tmem_relinquish();
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();     // no CTA multicasts into (or arrives on) a peer whose barriers are not initialised yet
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tmemP = C::kSplitP ? bars->tmem_p : tmem + C::kColP;
  // launched with programmatic stream serialization: the prologue above ran under the tail of the kernel in front (the
  // row-table prepass, and through it the previous attention call); nothing of global memory has been touched yet
  pdl_wait();

#ifdef CVD_WHATIF_NOSYNC
  // profiling build (make EXTRA=-DCVD_WHATIF_NOSYNC, results invalid): the softmax warps run their loop -- TMEM loads,
  // arithmetic, TMEM stores -- with every hand-off removed and no MMA / TMA issued: the cost of the softmax stream alone
  // in the production register / occupancy context
  constexpr bool kNoSync = true;
#else
  constexpr bool kNoSync = false;
#endif
  if (warp == 4 && kNoSync) {
    // one real Q K_0^T so that S holds scores (re-read by every "tile"); nothing else is issued
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::kQBytes + C::kKBytes);
#pragma unroll
      for (int kb = 0; kb < C::kKBlocks; ++kb) {
        tma_load_3d(sQ + kb * kBlockM * 128, &tmQ, kb * 64, h, r * N + qt * kBlockM, smem_u32(&bars->q_full));
        tma_load_3d(sK + kb * kTileN * 128, &kv.k[0], kb * 64, h, partner[r] * N, smem_u32(&bars->q_full));
      }
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars->q_full), 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < C::kKSteps; ++ks)
        umma_ss(tmem + C::kColS, umma_desc_k_sw128(sQ + (ks >> 2) * kBlockM * 128 + (ks & 3) * 32),
                umma_desc_k_sw128(sK + (ks >> 2) * kTileN * 128 + (ks & 3) * 32), kIdescQK, ks > 0);
      umma_commit(smem_u32(&bars->s_full));
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================== TMA producer + MMA issuer (one warp) =====================
    // Five warps per CTA: 3 CTAs = 15 warps per SM leave 128 registers per thread (18 warps: 96).  The whole warp walks
    // the loop (waits are warp-wide), one elected lane issues.  Ring stages are refilled by the issuer itself right
    // where their consumer MMA is known to have been issued long ago:
    //   K(j + stages) after Q K_j^T (k_empty), before Q K_{j+1}^T is issued;  V(j - 1 + stages) after P V_{j-1}.
    // Multicast build: every CTA arms its own "full" barrier for the bytes of the tile; the tile itself is fetched by
    // ONE CTA of the cluster (tile % CL) and delivered to all of them.  `refill` = the stage has been used before: the
    // fetching CTA first waits until the MMAs that read it have completed in EVERY CTA (CL multicast commit arrivals).
    auto load_k = [&](int lt, bool refill) {
      const int p = lt / C::kTiles, jt = lt % C::kTiles;
      const int pr = partner[p * R + r];
      const int src = pr / rows_per_src, rr = pr - src * rows_per_src;
      const int st = lt % C::kKStages;
      const bool mine = (CL == 1) || ((uint32_t)(lt % CL) == crank);
      if (refill && mine) mbar_wait(smem_u32(&bars->k_empty[st]), ((lt / C::kKStages) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->k_full[st]), C::kKBytes);
        if (mine) {
#pragma unroll
          for (int kb = 0; kb < C::kKBlocks; ++kb) {
            if (CL > 1)
              tma_load_3d_mc(sK + st * C::kKBytes + kb * kTileN * 128, &kv.k[src], kb * 64, h, rr * N + jt * kTileN,
                             smem_u32(&bars->k_full[st]), kMcMask);
            else
              tma_load_3d(sK + st * C::kKBytes + kb * kTileN * 128, &kv.k[src], kb * 64, h, rr * N + jt * kTileN,
                          smem_u32(&bars->k_full[st]));
          }
        }
      }
      __syncwarp();
    };
    auto load_v = [&](int lv, bool refill) {
      const int p = lv / C::kTiles, jt = lv % C::kTiles;
      const int pr = partner[p * R + r];
      const int src = pr / rows_per_src, rr = pr - src * rows_per_src;
      const int st = lv % C::kVStages;
      const bool mine = (CL == 1) || ((uint32_t)(lv % CL) == crank);
      if (refill && mine) mbar_wait(smem_u32(&bars->v_empty[st]), ((lv / C::kVStages) - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bars->v_full[st]), C::kVBytes);
        if (mine) {
          if (CL > 1)
            tma_load_2d_mc(sV + st * C::kVBytes, &kv.v[src], jt * kTileN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]),
                           kMcMask);
          else
            tma_load_2d(sV + st * C::kVBytes, &kv.v[src], jt * kTileN, (rr * H + h) * C::kDV, smem_u32(&bars->v_full[st]));
        }
      }
      __syncwarp();
    };
    // P in its own columns: QK(j+1) is issued as soon as the softmax warps have read S_j (s_free),
    // i.e. it runs under the softmax of tile j; PV(j) follows when P_j is complete.
    // P aliased on S (d=160): QK(j+1) has to follow PV(j) (tcgen05.mma executes in issue order).
    auto issue_qk = [&](int j) {
      const int st = j % C::kKStages;
      mbar_wait(smem_u32(&bars->k_full[st]), (j / C::kKStages) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < C::kKSteps; ++ks) {
          const uint32_t qoff = (ks >> 2) * kBlockM * 128 + (ks & 3) * 32;
          const uint32_t koff = (ks >> 2) * kTileN * 128 + (ks & 3) * 32;
          umma_ss(tmem + C::kColS, umma_desc_k_sw128(sQ + qoff), umma_desc_k_sw128(sK + st * C::kKBytes + koff),
                  kIdescQK, ks > 0);
        }
        umma_commit(smem_u32(&bars->s_full));
        if (CL > 1) umma_commit_mc(smem_u32(&bars->k_empty[st]), kMcMask);
        else umma_commit(smem_u32(&bars->k_empty[st]));
      }
      __syncwarp();
    };
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::kQBytes);
#pragma unroll
      for (int kb = 0; kb < C::kKBlocks; ++kb)
        tma_load_3d(sQ + kb * kBlockM * 128, &tmQ, kb * 64, h, r * N + qt * kBlockM, smem_u32(&bars->q_full));
    }
    __syncwarp();
    for (int lt = 0; lt < C::kKStages && lt < total; ++lt) load_k(lt, false);
    for (int lv = 0; lv < C::kVStages && lv < total; ++lv) load_v(lv, false);
    mbar_wait(smem_u32(&bars->q_full), 0);
    issue_qk(0);
    for (int j = 0; j < total; ++j) {
      const int st = j % C::kVStages;
      // DBG build: MMA-warp stamps, int64 [grid,32,8] behind the softmax stamps
      long long* ms = nullptr;
      if (DBG && j < 32 && lane == 0)
        ms = reinterpret_cast<long long*>(dbg + (size_t)gridDim.x * kBlockM * kTileN) + ((size_t)(gridDim.x + blockIdx.x) * 32 + j) * 8;
      if (DBG && ms) ms[0] = clock64();
      if (j + C::kKStages < total) load_k(j + C::kKStages, true);   // Q K_j^T has read its K stage: refill it
      if (!C::kAliasP && j + 1 < total) {
        mbar_wait(smem_u32(&bars->s_free), j & 1);
        if (DBG && ms) ms[1] = clock64();
        issue_qk(j + 1);
      }
      if (DBG && ms) ms[2] = clock64();
      if (j >= 1 && j - 1 + C::kVStages < total) load_v(j - 1 + C::kVStages, true);   // P V_{j-1} has read its V stage
      mbar_wait(smem_u32(&bars->v_full[st]), (j / C::kVStages) & 1);
      mbar_wait(smem_u32(&bars->p_full), j & 1);
      tc_fence_after();
      if (DBG && ms) ms[3] = clock64();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kTileN / 16; ++ks)
          umma_ts(tmem + C::kColO, tmemP + ks * 8, umma_desc_k_sw128(sV + st * C::kVBytes + ks * 32), kIdescPV,
                  (j > 0) || (ks > 0));
        umma_commit(smem_u32(&bars->pv_done));
        if (j == total - 1) umma_commit(smem_u32(&bars->o_done));
        if (CL > 1) umma_commit_mc(smem_u32(&bars->v_empty[st]), kMcMask);
        else umma_commit(smem_u32(&bars->v_empty[st]));
      }
      __syncwarp();
      if (C::kAliasP && j + 1 < total) issue_qk(j + 1);
      if (DBG && ms) ms[4] = clock64();
    }
    if (CL > 1) {
      // A CTA may only leave once no peer can still signal it: the last use of every stage must have collected the
      // commit arrivals of ALL CTAs of the cluster (their final MMAs on that stage have completed).
#pragma unroll
      for (int st = 0; st < C::kKStages; ++st)
        if (st < total) mbar_wait(smem_u32(&bars->k_empty[st]), (((total - 1 - st) / C::kKStages)) & 1);
#pragma unroll
      for (int st = 0; st < C::kVStages; ++st)
        if (st < total) mbar_wait(smem_u32(&bars->v_empty[st]), (((total - 1 - st) / C::kVStages)) & 1);
    }
  } else {
    // =============================== softmax warps ==============================
    const int tid = threadIdx.x;                       // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t tS = tmem + lane_off + C::kColS, tO = tmem + lane_off + C::kColO, tP = tmemP + lane_off;
    int n = qt * kBlockM + tid;
    const bool row_valid = n < N;
    if (!row_valid) n = N - 1;
    const float xq = pixel_centre(n % FEAT, step), yq = pixel_centre(n / FEAT, step);
    const float off = (step - 1.0f) * 0.5f;
    const float scale_log2 = kLog2e * rsqrtf((float)D);

    // bias geometry of this thread's query pixel, base-2 units (already scaled by w_r * log2e / |l_q|):
    //   t(key) = As * col + Bs * row + base,   bias = min(band' - |t|, 0) = min(band' - t, band' + t, 0)
    float As = 0.f, Bs = 0.f, base = 0.f, bandp = 0.f;
    float m_used = 0.f;                   // running reference maximum (set by the first half tile)
    float bias_ub = 0.f;                  // upper bound of this query's bias over ALL keys of the first partner
    const uint64_t scale2 = pack2(scale_log2, scale_log2);

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
      if (DBG && tid == 0 && j < 32)
        stamps = reinterpret_cast<long long*>(dbg + (size_t)gridDim.x * kBlockM * kTileN) + ((size_t)blockIdx.x * 32 + j) * 8;
      if (DBG && stamps) stamps[0] = clock64();
      if (!kNoSync || j == 0) mbar_wait(smem_u32(&bars->s_full), j & 1);
      tc_fence_after();
      if (DBG && stamps) stamps[1] = clock64();

      uint32_t preg[kTileN / 2];
      if constexpr (PIPE) {
        // Fused form: the tile is walked in four 16-key quarters and every score goes bias -> FFMA2 -> MUFU.EX2 in ONE
        // straight-line block, so each ex2 depends on the arithmetic of its own element and ptxas interleaves the MUFU
        // stream with the FADD2 / FMNMX3 / FFMA2 stream (a warp keeps both pipes busy by itself; with the vote between
        // the two phases it ran them back to back: 15 vs 9 cycles per warp-column).  The reference-maximum vote FOLLOWS
        // the exponentials: in the rare case that a score exceeds the reference by more than 2^16 the quarter's
        // exponentials are recomputed against the new reference from the x values still in registers (exact for any
        // magnitude), and everything older is rescaled as before.  The second half of S is fetched under block 0.
        constexpr int kQ = 16;
        uint32_t s0[kHalfN], s1[kHalfN];
        uint32_t sq[kQ];
        if (!C::kFour) {
          tmem_ld_x32(tS, s0);
          tmem_ld_wait();
        }
        if (DBG && j == 0 && !C::kFour) {
          float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kTileN;
#pragma unroll
          for (int c = 0; c < kHalfN; ++c) o[c] = __uint_as_float(s0[c]);
        }
        const uint64_t step1 = pack2(-2.f * As, -2.f * As), step2 = pack2(2.f * As, 2.f * As);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t* cur = C::kFour ? sq : ((q < 2 ? s0 : s1) + (q & 1) * kQ);   // scores -> x of quarter q (in place)
          if (C::kFour) {                                                    // 96-register build: 16 scores at a time
            tmem_ld_x16(tS + q * kQ, sq);
            tmem_ld_wait();
            if (DBG && j == 0) {
              float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kTileN + q * kQ;
#pragma unroll
              for (int c = 0; c < kQ; ++c) o[c] = __uint_as_float(sq[c]);
            }
          }
          if (q == 0 && !C::kFour) tmem_ld_x32(tS + kHalfN, s1);             // lands under this block
          const float bm = bandp - m_used, negm = -m_used;
          float xmax = -INFINITY;
          uint64_t a1 = 0, a2 = 0;
#pragma unroll
          for (int c = 0; c < kQ; c += 2) {
            const int cc = q * kQ + c;           // key index inside the 64-key tile
            if (cc % FEAT == 0 || c == 0) {      // start of a key-grid row (or of this quarter)
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
            unpack2(fma2(pack2(__uint_as_float(cur[c]), __uint_as_float(cur[c + 1])), scale2, pack2(v0, v1)), x0, x1);
            if (N < kTileN && cc >= N) x0 = x1 = -INFINITY;   // keys past the row (N = 16): not in the softmax
            xmax = max3(xmax, x0, x1);
            preg[q * (kQ / 2) + (c >> 1)] = pack_bf16(ex2(x0), ex2(x1));
            cur[c] = __float_as_uint(x0);
            cur[c + 1] = __float_as_uint(x1);
            a1 = add2(a1, step1);
            a2 = add2(a2, step2);
          }
          if (q == 0 && !C::kFour) {
            tmem_ld_wait_dep32(s1);
            if (!C::kAliasP && !kNoSync) {      // S_j is in registers: the tensor core may start Q K_{j+1}^T
              tc_fence_before();
              mbar_arrive(smem_u32(&bars->s_free));
            }
            if (DBG && j == 0) {
              float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kTileN + kHalfN;
#pragma unroll
              for (int c = 0; c < kHalfN; ++c) o[c] = __uint_as_float(s1[c]);
            }
          }
          // lazy maximum update (always on the very first quarter)
          const bool first = (j == 0 && q == 0);
          const bool need = first || (xmax > kRescaleThreshold);
          if (__any_sync(0xffffffffu, need)) {
            if (!first) {
              const float f = need ? ex2(-xmax) : 1.0f;
              if (j > 0) {
                if (!kNoSync) mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
                tc_fence_after();
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
              // the earlier quarters of P_j were computed against the old reference
#pragma unroll
              for (int i = 0; i < q * (kQ / 2); ++i)
                preg[i] = pack_bf16(__uint_as_float(preg[i] << 16) * f, __uint_as_float(preg[i] & 0xffff0000u) * f);
            }
            // this quarter: exponentials again, against the new reference
            const float shift = need ? (first ? fmaxf(xmax, bias_ub) : xmax) : 0.f;
            m_used += shift;
#pragma unroll
            for (int c = 0; c < kQ; c += 2)
              preg[q * (kQ / 2) + (c >> 1)] =
                  pack_bf16(ex2(__uint_as_float(cur[c]) - shift), ex2(__uint_as_float(cur[c + 1]) - shift));
          }
          if (DBG && stamps && (q & 1)) stamps[2 + (q >> 1)] = clock64();
        }
      } else {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t sreg[kHalfN];
        tmem_ld_x32(tS + hh * kHalfN, sreg);
        tmem_ld_wait();
        if (hh == 1 && !C::kAliasP) {      // S_j is in registers: the tensor core may start Q K_{j+1}^T
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->s_free));
        }
        if (DBG && j == 0) {
          float* o = dbg + ((size_t)blockIdx.x * kBlockM + tid) * kTileN + hh * kHalfN;
#pragma unroll
          for (int c = 0; c < kHalfN; ++c) o[c] = __uint_as_float(sreg[c]);
        }
        // x = qk * log2e/sqrt(d) + bias - m_used, evaluated against the CURRENT reference maximum so
        // the subtraction is free: min(band'-m - t, band'-m + t, -m) is one FMNMX3 per key, the two
        // linear terms advance by one packed FADD2 per key pair, the score is one packed FFMA2.
        const float bm = bandp - m_used, negm = -m_used;
        const uint64_t step1 = pack2(-2.f * As, -2.f * As), step2 = pack2(2.f * As, 2.f * As);
        float xmax = -INFINITY;
        uint64_t a1 = 0, a2 = 0;
#pragma unroll
        for (int c = 0; c < kHalfN; c += 2) {
          const int cc = hh * kHalfN + c;      // key index inside the 64-key tile
          if (cc % FEAT == 0 || c == 0) {      // start of a key-grid row (or of this half)
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
          unpack2(fma2(pack2(__uint_as_float(sreg[c]), __uint_as_float(sreg[c + 1])), scale2, pack2(v0, v1)), x0, x1);
          if (N < kTileN && cc >= N) x0 = x1 = -INFINITY;   // keys past the row (N = 16): not in the softmax
          xmax = max3(xmax, x0, x1);
          sreg[c] = __float_as_uint(x0);
          sreg[c + 1] = __float_as_uint(x1);
          a1 = add2(a1, step1);
          a2 = add2(a2, step2);
        }

        // lazy maximum update (always on the very first half tile)
        const bool first = (j == 0 && hh == 0);
        const bool need = first || (xmax > kRescaleThreshold);
        if (__any_sync(0xffffffffu, need)) {
          if (!first) {
            const float f = need ? ex2(-xmax) : 1.0f;
            if (j > 0) {
              // O holds tiles 0..j-1: PV(j-1) must have landed.  (s_full(j) implies PV(j-2) has completed,
              // so this thread is at most one phase behind pv_done.)
              mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
              tc_fence_after();
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
            if (hh == 1) {   // the first half of P_j (still in registers) was computed against the old reference
#pragma unroll
              for (int i = 0; i < kHalfN / 2; ++i)
                preg[i] = pack_bf16(__uint_as_float(preg[i] << 16) * f, __uint_as_float(preg[i] & 0xffff0000u) * f);
            }
          }
          if (need) {
            // First half tile: anchor the reference at the largest score this row can still reach through
            // the bias (the epipolar band may lie many key rows ahead), so that approaching the band does
            // not trigger a rescale every other tile.  Any reference >= the running maximum is valid.
            const float shift = first ? fmaxf(xmax, bias_ub) : xmax;
            m_used += shift;
#pragma unroll
            for (int c = 0; c < kHalfN; ++c) sreg[c] = __float_as_uint(__uint_as_float(sreg[c]) - shift);
          }
        }
#pragma unroll
        for (int c = 0; c < kHalfN; c += 2)
          preg[hh * (kHalfN / 2) + (c >> 1)] = pack_bf16(ex2(__uint_as_float(sreg[c])), ex2(__uint_as_float(sreg[c + 1])));
        if (DBG && stamps) stamps[2 + hh] = clock64() + (long long)(preg[hh * 16] & 1u);
      }
      }
      if (j > 0 && !kNoSync) {   // the P columns are free once PV(j-1) has read them (aliased mode: implied by s_full(j))
        mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
        tc_fence_after();
      }
      if (DBG && stamps) stamps[4] = clock64();
      tmem_st_x32(tP, preg);
      tmem_st_wait();
      tc_fence_before();
      if (!kNoSync) mbar_arrive(smem_u32(&bars->p_full));
      if (DBG && stamps) stamps[5] = clock64();
    }

    // epilogue: O[:, 0:D] / O[:, D]  ->  out[r, n, h*D : (h+1)*D]
    if (!kNoSync) mbar_wait(smem_u32(&bars->o_done), 0);   // committed once, after the last PV
    tc_fence_after();
    uint32_t lraw;
    tmem_ld_x1(tO + D, lraw);
    tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(lraw);
    // base-2 log-sum-exp of the row's logits (scores * log2e/sqrt(d) + bias * log2e): what the backward pass needs
    if (lse != nullptr && row_valid) lse[((size_t)r * H + h) * N + n] = m_used + log2f(__uint_as_float(lraw));
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
      if (row_valid) *reinterpret_cast<uint4*>(orow + c0) = pk;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem);
    if (C::kSplitP) tmem_dealloc<32>(bars->tmem_p);
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

// Encoded descriptors are cached per (base pointer, shape, box): a UNet forward issues the same 40 (q, k, v^T) triples
// every denoising step (PyTorch's caching allocator hands the same blocks back), so steady-state launches do no driver
// calls.  A descriptor only encodes address and geometry, so a stale entry for a recycled pointer is still correct.
struct MapKey {
  const void* base;
  uint64_t a, b, c;
  uint32_t box_rows, kind;
  bool operator<(const MapKey& o) const {
    return std::tie(base, a, b, c, box_rows, kind) < std::tie(o.base, o.a, o.b, o.c, o.box_rows, o.kind);
  }
};
static std::mutex g_map_mu;
static std::map<MapKey, CUtensorMap> g_maps;

template <typename Encode>
static int cached_map(CUtensorMap* m, const MapKey& key, Encode&& encode) {
  {
    std::lock_guard<std::mutex> lock(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *m = it->second;
      return CVD_OK;
    }
  }
  const int rc = encode();
  if (rc != CVD_OK) return rc;
  std::lock_guard<std::mutex> lock(g_map_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps[key] = *m;
  return CVD_OK;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous), box [box_rows, 64 cols], 128-byte swizzle
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return cached_map(m, MapKey{base, rows, cols, 0, box_rows, 2}, [&]() -> int {
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
  });
}

// bf16 projections [rows, heads * d] seen as 3-D [rows, heads, d]; box [box_rows, 1 head, 64 columns], 128-byte
// swizzle.  The 64-wide box is wider than a head for d < 64 (and its last box for d % 64 != 0): the columns past
// d are out of bounds of the innermost dimension and are filled with zeros by the TMA unit, so the K padding of
// the MMA needs no patching and no bytes of the neighbouring head are fetched.
static int make_map_heads(CUtensorMap* m, const void* base, uint64_t rows, uint64_t heads, uint64_t d, uint32_t box_rows) {
  return cached_map(m, MapKey{base, rows, heads, d, box_rows, 3}, [&]() -> int {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CVD_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[3] = {d, heads, rows};
    const cuuint64_t strides[2] = {d * sizeof(__nv_bfloat16), heads * d * sizeof(__nv_bfloat16)};
    const cuuint32_t box[3] = {64, 1, box_rows};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
      return fail(CVD_E_CUDA, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d (rows=%llu heads=%llu d=%llu box_rows=%u)",
                  (int)rc, (unsigned long long)rows, (unsigned long long)heads, (unsigned long long)d, box_rows);
    return CVD_OK;
  });
}

template <int D, int FEAT, int CL>
static int launch(const void* q, const KvSrc& src, const float* rowp, const int* partner, void* out, int R, int H,
                  int n_partners, float S, float* dbg, float* lse, cudaStream_t st) {
  using C = Cfg<D, FEAT, CL>;
  constexpr int N = FEAT * FEAT;
  CUtensorMap tmQ;
  KvMaps kv;
  int rc;
  if ((rc = make_map_heads(&tmQ, q, (uint64_t)R * N, (uint64_t)H, (uint64_t)D, kBlockM)) != CVD_OK) return rc;
  for (int i = 0; i < kMaxSrc; ++i) {
    const int j = i < src.n_src ? i : 0;          // unused slots repeat source 0 (never addressed)
    if (i < src.n_src) {
      if ((rc = make_map_heads(&kv.k[i], src.k[j], (uint64_t)src.rows_per_src * N, (uint64_t)H, (uint64_t)D, kTileN)) != CVD_OK) return rc;
      if ((rc = make_map(&kv.v[i], src.v[j], (uint64_t)src.rows_per_src * H * C::kDV, (uint64_t)N, C::kDV)) != CVD_OK) return rc;
    } else {
      kv.k[i] = kv.k[0];
      kv.v[i] = kv.v[0];
    }
  }
  constexpr bool kHasDebug = (D == 40 && FEAT == 32 && CL == 1);   // the S-dump variant is built for the bench shape only
  CVD_REQUIRE(dbg == nullptr || kHasDebug, "cvd_epi_attn_fwd_debug: only (head_dim=40, feat=32) has a debug build");
  // CVD_ATTN_PIPE=0 selects the un-pipelined softmax loop at the bench shape (A/B measurements only)
  static int pipe = -1;
  if (pipe < 0) {
    const char* e = getenv("CVD_ATTN_PIPE");
    pipe = (e != nullptr && e[0] == '0' && kHasDebug) ? 0 : 1;
  }
  auto kern = (dbg != nullptr && kHasDebug) ? epi_attn_fwd_kernel<D, FEAT, kHasDebug, true, CL>
                                            : (pipe ? epi_attn_fwd_kernel<D, FEAT, false, true, CL>
                                                    : epi_attn_fwd_kernel<D, FEAT, false, !kHasDebug, CL>);
  CVD_CUDA(ensure_dynamic_smem((const void*)kern, C::kSmemBytes));
  constexpr int kQTiles = (N + kBlockM - 1) / kBlockM;
  static_assert(kQTiles % CL == 0, "a cluster holds query tiles of ONE (row, head)");
  const long long grid = (long long)R * H * kQTiles;
  CVD_REQUIRE(grid > 0 && grid < (1ll << 31), "cvd_epi_attn_fwd: grid of %lld CTAs", grid);
  CVD_CUDA(launch_pdl_cluster(kern, dim3((unsigned)grid), dim3(kThreads), (size_t)C::kSmemBytes, st, (unsigned)CL, tmQ, kv, rowp,
                              partner, (__nv_bfloat16*)out, R, H, n_partners, src.rows_per_src, S / (float)FEAT, dbg, lse));
  CVD_LAUNCH_CHECK("epi_attn_fwd_kernel");
  return CVD_OK;
}

// Cluster-multicast kernels: used for K / V^T in LOCAL memory only.  MEASURED (round 2, 2 x B200, driver 580.159):
// cp.async.bulk.tensor ... .multicast::cluster with a source in a PEER GPU's memory (NVLink mapping) never completes and
// leaves the GPU unresponsive (the time-bounded mbarrier waits never get to trap), while the same kernel on local
// copies of the same rows is bit-exact and unicast TMA reads of peer memory work.  So the cluster kernels
// are opt-in (CVD_ATTN_CLUSTER=1 in the environment, read at every call) and, as a second line of defence, refused unless
// every source is memory of the current device.
static bool cluster_requested() {
  const char* e = getenv("CVD_ATTN_CLUSTER");
  return e != nullptr && e[0] == '1';
}

static bool all_sources_local(const KvSrc& src) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  for (int i = 0; i < src.n_src; ++i) {
    const void* ptrs[2] = {src.k[i], src.v[i]};
    for (const void* p : ptrs) {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      if (a.type != cudaMemoryTypeDevice || a.device != dev) return false;
    }
  }
  return true;
}

static int dispatch(const void* q, const KvSrc& src, const float* rowp, const int* partner, void* out, int R, int N, int heads,
                    int head_dim, int n_partners, int feat, float S, float* dbg, float* lse, cudaStream_t st) {
  CVD_REQUIRE(q && partner && out, "cvd_epi_attn_fwd: null pointer");
  CVD_REQUIRE(src.n_src >= 1 && src.n_src <= kMaxSrc && src.rows_per_src > 0, "cvd_epi_attn_fwd: %d K/V sources of %d rows", src.n_src,
              src.rows_per_src);
  CVD_REQUIRE(R > 0 && heads > 0 && n_partners >= 1, "cvd_epi_attn_fwd: R=%d heads=%d n_partners=%d", R, heads, n_partners);
  CVD_REQUIRE(N == feat * feat, "cvd_epi_attn_fwd: N=%d is not feat^2 (feat=%d); only square latents are supported", N, feat);
  uintptr_t bits = (uintptr_t)q | (uintptr_t)out;
  for (int i = 0; i < src.n_src; ++i) {
    CVD_REQUIRE(src.k[i] && src.v[i], "cvd_epi_attn_fwd: null K / V^T pointer (source %d)", i);
    bits |= (uintptr_t)src.k[i] | (uintptr_t)src.v[i];
  }
  CVD_REQUIRE((bits & 15) == 0, "cvd_epi_attn_fwd: q/k/vt/out must be 16-byte aligned");
  if (dbg == nullptr && feat <= 4) {   // 4x4 latent grid: warp-level kernel (a 128x64 tcgen05 tile would be 97 % padding; 6.5 vs 11 us)
    const int rc = epi_attn_small_dispatch(q, src, rowp, partner, out, lse, R, heads, head_dim, n_partners, feat, S, st);
    if (rc != -1000) return rc;
  }
  const bool mc = dbg == nullptr && cluster_requested() && all_sources_local(src);
#define CVD_CASE(D_, F_, CL_)                                                                                       \
  if (head_dim == D_ && feat == F_) {                                                                               \
    if (mc && CL_ > 1) return launch<D_, F_, CL_>(q, src, rowp, partner, out, R, heads, n_partners, S, dbg, lse, st); \
    return launch<D_, F_, 1>(q, src, rowp, partner, out, R, heads, n_partners, S, dbg, lse, st);                     \
  }
  CVD_CASE(40, 32, 8)
  CVD_CASE(40, 64, 8)
  CVD_CASE(80, 16, 2)
  CVD_CASE(80, 32, 8)
  CVD_CASE(160, 16, 2)
  CVD_CASE(40, 16, 2)
  CVD_CASE(160, 8, 1)
  CVD_CASE(80, 8, 1)
  CVD_CASE(40, 8, 1)
#undef CVD_CASE
  return fail(CVD_E_INVALID, "cvd_epi_attn_fwd: (head_dim=%d, feat=%d) is not an instantiated configuration", head_dim, feat);
}

static KvSrc one_source(const void* k, const void* vt, int R_kv) {
  KvSrc s = {};
  s.k[0] = k;
  s.v[0] = vt;
  s.n_src = 1;
  s.rows_per_src = R_kv;
  return s;
}

}  // namespace sm100
}  // namespace cvd

extern "C" int cvd_epi_attn_fwd(const void* q, const void* k, const void* vt, const float* rowp, const int* partner,
                                void* out, int R, int R_kv, int N, int heads, int head_dim, int n_partners, int feat,
                                float F_mat_size, cvd_stream_t stream) {
  CVD_REQUIRE(R_kv > 0, "cvd_epi_attn_fwd: R_kv=%d", R_kv);
  return cvd::sm100::dispatch(q, cvd::sm100::one_source(k, vt, R_kv), rowp, partner, out, R, N, heads, head_dim, n_partners, feat,
                              F_mat_size, nullptr, nullptr, (cudaStream_t)stream);
}

// View-sharded form: K / V^T of the partner rows live in n_src buffers of rows_per_src rows each (the symmetric-memory
// buffers of the GPUs of one NVSwitch box, passed as peer-mapped device pointers, own rank included); `partner` holds
// GLOBAL rows.  The kernel reads the partner tiles straight out of the owning GPU's HBM (unicast TMA).
extern "C" int cvd_epi_attn_fwd_sharded(const void* q, const void* const* k_srcs, const void* const* vt_srcs, int n_src,
                                        int rows_per_src, const float* rowp, const int* partner, void* out, int R, int N,
                                        int heads, int head_dim, int n_partners, int feat, float F_mat_size,
                                        cvd_stream_t stream) {
  CVD_REQUIRE(k_srcs && vt_srcs && n_src >= 1 && n_src <= cvd::kMaxKvSrc, "cvd_epi_attn_fwd_sharded: n_src=%d (1..%d)", n_src,
              cvd::kMaxKvSrc);
  cvd::KvSrc s = {};
  for (int i = 0; i < n_src; ++i) {
    s.k[i] = k_srcs[i];
    s.v[i] = vt_srcs[i];
  }
  s.n_src = n_src;
  s.rows_per_src = rows_per_src;
  return cvd::sm100::dispatch(q, s, rowp, partner, out, R, N, heads, head_dim, n_partners, feat, F_mat_size, nullptr, nullptr,
                              (cudaStream_t)stream);
}

// Same as cvd_epi_attn_fwd, additionally writing the base-2 log-sum-exp of every (row, head, query) -- the statistic the
// backward pass (cvd_epi_attn_bwd) recomputes the probabilities from.  lse: fp32 [R, heads, N].
extern "C" int cvd_epi_attn_fwd_lse(const void* q, const void* k, const void* vt, const float* rowp, const int* partner,
                                    void* out, float* lse, int R, int R_kv, int N, int heads, int head_dim, int n_partners,
                                    int feat, float F_mat_size, cvd_stream_t stream) {
  CVD_REQUIRE(lse != nullptr && R_kv > 0, "cvd_epi_attn_fwd_lse: null lse / R_kv=%d", R_kv);
  return cvd::sm100::dispatch(q, cvd::sm100::one_source(k, vt, R_kv), rowp, partner, out, R, N, heads, head_dim, n_partners, feat,
                              F_mat_size, nullptr, lse, (cudaStream_t)stream);
}

// Bring-up aid: additionally dumps the raw Q K^T accumulator of key tile 0 of every CTA
// (dbg_s: fp32 [grid, 128, 64]) so the first MMA can be verified independently of the rest.
extern "C" int cvd_epi_attn_fwd_debug(const void* q, const void* k, const void* vt, const float* rowp, const int* partner,
                                      void* out, float* dbg_s, int R, int R_kv, int N, int heads, int head_dim,
                                      int n_partners, int feat, float F_mat_size, cvd_stream_t stream) {
  CVD_REQUIRE(R_kv > 0, "cvd_epi_attn_fwd_debug: R_kv=%d", R_kv);
  return cvd::sm100::dispatch(q, cvd::sm100::one_source(k, vt, R_kv), rowp, partner, out, R, N, heads, head_dim, n_partners, feat,
                              F_mat_size, dbg_s, nullptr, (cudaStream_t)stream);
}
