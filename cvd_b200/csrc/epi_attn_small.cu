// Epipolar cross-view attention for the smallest latent grids (N = feat^2 <= 64 keys per partner;
// instantiated and dispatched for the 4x4 level of the UNet; d = 160 / 80 / 40).  Same contract as the tcgen05 kernel
// (csrc/epi_attn_sm100.cu; reference: animatediff/models/epi_module.py:293-329,395-407 and
// animatediff/models/attention_processor.py:573-619), different machine mapping:
//
// At these sizes one (row, head) problem is a 64x64 (or 16x16) score matrix -- a single 128-row
// tcgen05 tile would be 50-88 % padding, and a CTA's life would be TMEM allocation, barrier set-up
// and TMA round trips (measured at N = 16: 11 us per launch on the tcgen05 kernel, 6.5 us here; at
// N = 64, d = 160 the tcgen05 kernel wins, 14.5 vs 31 us, because this kernel's fragment loads are
// latency-bound at 182 registers per thread -- the dispatcher picks per shape).  The work is bandwidth/latency
// shaped, so it is mapped to warps instead: one warp owns 16 query rows of one (row, head), keeps
// the whole score strip S[16, M] in registers (warp-level mma.sync m16n8k16, bf16 in / fp32
// accumulate), applies the in-register epipolar bias, takes an exact row softmax (no online
// rescale needed: all keys are present), reuses the accumulator registers as the A operand of
// P V, and writes bf16 rows.  No shared memory, no block-level synchronisation; q is read once,
// k / v^T fragments come through L1 (four warps of a CTA share the same k / v^T at N = 64).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cvd {
namespace small {

using ptx::ex2;
using ptx::pack_bf16;

constexpr int kWarps = 4;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t ldg_u32(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }

// One warp task = 16 query rows [m0, m0+16) of (row r, head h).
template <int D, int FEAT, int KP>   // KP: partners whose keys are concatenated (1 or 2)
__global__ void __launch_bounds__(kWarps * 32)
epi_attn_small_kernel(const __nv_bfloat16* __restrict__ q, const KvSrc kv, const float* __restrict__ rowp,
                      const int* __restrict__ partner, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int R, int H,
                      float step) {
  constexpr int N = FEAT * FEAT;
  constexpr int kMBlocks = N / 16;                 // 16-row query blocks per (row, head)
  constexpr int kKSteps = (D + 15) / 16;           // k steps of Q K^T
  constexpr int kNT = N / 8;                       // 8-key tiles per partner
  constexpr int kDV = (D / 16 + 1) * 16;           // rows of a v^T head block (d values, ones row, padding)
  constexpr int kDT = D / 8;                       // 8-wide output tiles
  static_assert(N % 16 == 0 && N <= 64 && D % 8 == 0, "small-grid kernel: N in {16, 64}");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_trigger();      // programmatic dependent launch (common.cuh): the successor may start under this grid's tail,
  pdl_wait();         // and this grid was allowed to start under its predecessor's
  const long long task = (long long)blockIdx.x * kWarps + warp;
  if (task >= (long long)R * H * kMBlocks) return;
  const int mb = (int)(task % kMBlocks), h = (int)((task / kMBlocks) % H), r = (int)(task / ((long long)kMBlocks * H));
  const int C = H * D;
  const int g = lane >> 2, t = lane & 3;           // fragment row group / column pair
  const int row0 = mb * 16 + g, row1 = row0 + 8;   // the two query rows this thread holds

  // ---- Q fragments (A operand, row-major): dims >= D are zero (head stride has no padding)
  uint32_t qa[kKSteps][4];
  {
    const __nv_bfloat16* q0 = q + ((size_t)r * N + row0) * C + h * D;
    const __nv_bfloat16* q1 = q + ((size_t)r * N + row1) * C + h * D;
#pragma unroll
    for (int ks = 0; ks < kKSteps; ++ks) {
      const int d0 = ks * 16 + t * 2, d1 = d0 + 8;
      qa[ks][0] = d0 < D ? ldg_u32(q0 + d0) : 0u;
      qa[ks][1] = d0 < D ? ldg_u32(q1 + d0) : 0u;
      qa[ks][2] = d1 < D ? ldg_u32(q0 + d1) : 0u;
      qa[ks][3] = d1 < D ? ldg_u32(q1 + d1) : 0u;
    }
  }

  // ---- S = Q K^T over the keys of all partners, scaled to base-2 logits with the epipolar bias
  float s[KP][kNT][4];
  const float scale_log2 = kLog2e * rsqrtf((float)D);
  const float off = (step - 1.0f) * 0.5f;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int p = 0; p < KP; ++p) {
    const int pr = partner[p * R + r];
    const int src = pr / kv.rows_per_src, rr = pr - src * kv.rows_per_src;
    const __nv_bfloat16* kb = (const __nv_bfloat16*)kv.k[src] + (size_t)rr * N * C + h * D;
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      s[p][nt][0] = s[p][nt][1] = s[p][nt][2] = s[p][nt][3] = 0.f;
      const __nv_bfloat16* krow = kb + (size_t)(nt * 8 + g) * C;   // B operand: key = nt*8 + g, dims (t*2, t*2+1) (+8)
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks) {
        const int d0 = ks * 16 + t * 2, d1 = d0 + 8;
        const uint32_t b0 = d0 < D ? ldg_u32(krow + d0) : 0u;
        const uint32_t b1 = d1 < D ? ldg_u32(krow + d1) : 0u;
        mma_bf16_16816(s[p][nt], qa[ks], b0, b1);
      }
    }
    // bias of this partner for the thread's two query pixels (base-2 units, see epi_attn_sm100.cu)
    float As[2] = {0.f, 0.f}, Bs[2] = {0.f, 0.f}, base[2] = {0.f, 0.f}, bandp = 0.f;
    if (rowp != nullptr) {
      const float* P = rowp + (size_t)(p * R + r) * CVD_ROWP_STRIDE;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n = i == 0 ? row0 : row1;
        const float xq = pixel_centre(n % FEAT, step), yq = pixel_centre(n / FEAT, step);
        const float a = fmaf(P[1], yq, P[0] * xq) + P[2];
        const float b = fmaf(P[4], yq, P[3] * xq) + P[5];
        const float c = fmaf(P[7], yq, P[6] * xq) + P[8];
        const float k1 = P[10] * kLog2e / (sqrtf(a * a + b * b) + 1e-6f);
        As[i] = a * k1 * step;
        Bs[i] = b * k1 * step;
        base[i] = fmaf(a * k1, off, fmaf(b * k1, off, c * k1));
      }
      bandp = P[9] * P[10] * kLog2e;
    }
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + t * 2 + (e & 1);
        const int i = e >> 1;                       // 0: row0, 1: row1
        float x = s[p][nt][e] * scale_log2;
        if (rowp != nullptr) {
          const float tt = fmaf(As[i], (float)(key % FEAT), fmaf(Bs[i], (float)(key / FEAT), base[i]));
          float bias = fminf(bandp - fabsf(tt), 0.f);
          if (!(fabsf(bias) <= 3.0e38f)) bias = 0.f;   // NaN / inf geometry -> 0 (epi_module.py:404-407)
          x += bias;
        }
        s[p][nt][e] = x;
        if (i == 0) mx0 = fmaxf(mx0, x); else mx1 = fmaxf(mx1, x);
      }
    }
  }
  // exact row maximum / sum over the 4 threads that share a row
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t pa[KP][kNT][2];                          // P as packed bf16 pairs: [..][0] row0, [..][1] row1
#pragma unroll
  for (int p = 0; p < KP; ++p)
#pragma unroll
    for (int nt = 0; nt < kNT; ++nt) {
      const float e0 = ex2(s[p][nt][0] - mx0), e1 = ex2(s[p][nt][1] - mx0);
      const float e2 = ex2(s[p][nt][2] - mx1), e3 = ex2(s[p][nt][3] - mx1);
      sum0 += e0 + e1;
      sum1 += e2 + e3;
      pa[p][nt][0] = pack_bf16(e0, e1);
      pa[p][nt][1] = pack_bf16(e2, e3);
    }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  if (lse != nullptr && t == 0) {      // base-2 log-sum-exp per query row (backward pass)
    lse[((size_t)r * H + h) * N + row0] = mx0 + log2f(sum0);
    lse[((size_t)r * H + h) * N + row1] = mx1 + log2f(sum1);
  }

  // ---- O = P V: the score accumulators are already laid out as the A operand; B from v^T [dim][key]
  __nv_bfloat16* o0 = out + ((size_t)r * N + row0) * C + h * D;
  __nv_bfloat16* o1 = out + ((size_t)r * N + row1) * C + h * D;
#pragma unroll 2
  for (int dt = 0; dt < kDT; ++dt) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < KP; ++p) {
      const int pr = partner[p * R + r];
      const int src = pr / kv.rows_per_src, rr = pr - src * kv.rows_per_src;
      const __nv_bfloat16* vrow = (const __nv_bfloat16*)kv.v[src] + ((size_t)(rr * H + h) * kDV + dt * 8 + g) * N;   // B operand: dim = dt*8 + g, keys (t*2, +1) (+8)
#pragma unroll
      for (int kk = 0; kk < N / 16; ++kk) {
        const uint32_t a[4] = {pa[p][2 * kk][0], pa[p][2 * kk][1], pa[p][2 * kk + 1][0], pa[p][2 * kk + 1][1]};
        const uint32_t b0 = ldg_u32(vrow + kk * 16 + t * 2), b1 = ldg_u32(vrow + kk * 16 + t * 2 + 8);
        mma_bf16_16816(o, a, b0, b1);
      }
    }
    *reinterpret_cast<uint32_t*>(o0 + dt * 8 + t * 2) = pack_bf16(o[0] * inv0, o[1] * inv0);
    *reinterpret_cast<uint32_t*>(o1 + dt * 8 + t * 2) = pack_bf16(o[2] * inv1, o[3] * inv1);
  }
}

template <int D, int FEAT>
static int launch(const void* q, const KvSrc& kv, const float* rowp, const int* partner, void* out, float* lse, int R, int H,
                  int n_partners, float S, cudaStream_t st) {
  constexpr int N = FEAT * FEAT;
  const long long tasks = (long long)R * H * (N / 16);
  const long long grid = (tasks + kWarps - 1) / kWarps;
  CVD_REQUIRE(grid > 0 && grid < (1ll << 31), "cvd_epi_attn_fwd: grid of %lld CTAs", grid);
  CVD_REQUIRE(n_partners == 1 || n_partners == 2, "cvd_epi_attn_fwd: %d partners at feat=%d (1 or 2 are instantiated)", n_partners, FEAT);
  const float step = S / (float)FEAT;
  if (n_partners == 1)
    CVD_CUDA(launch_pdl(epi_attn_small_kernel<D, FEAT, 1>, dim3((unsigned)grid), dim3(kWarps * 32), 0, st, (const __nv_bfloat16*)q,
                        kv, rowp, partner, (__nv_bfloat16*)out, lse, R, H, step));
  else
    CVD_CUDA(launch_pdl(epi_attn_small_kernel<D, FEAT, 2>, dim3((unsigned)grid), dim3(kWarps * 32), 0, st, (const __nv_bfloat16*)q,
                        kv, rowp, partner, (__nv_bfloat16*)out, lse, R, H, step));
  CVD_LAUNCH_CHECK("epi_attn_small_kernel");
  return CVD_OK;
}

}  // namespace small

// Returns CVD_OK / an error if (head_dim, feat) is a small-grid configuration, and -1000 if it is not
// (the caller then uses the tcgen05 kernel).
int epi_attn_small_dispatch(const void* q, const KvSrc& kv, const float* rowp, const int* partner, void* out, float* lse, int R,
                            int heads, int head_dim, int n_partners, int feat, float S, cudaStream_t st) {
#define CVD_CASE(D_, F_) \
  if (head_dim == D_ && feat == F_) return small::launch<D_, F_>(q, kv, rowp, partner, out, lse, R, heads, n_partners, S, st);
  CVD_CASE(160, 4)
  CVD_CASE(80, 4)
  CVD_CASE(40, 4)
#undef CVD_CASE
  return -1000;
}

}  // namespace cvd
