// Plain CUDA-core fp32 epipolar cross-view attention: the on-device cross-check of the
// tcgen05 kernel.  Follows EpiAttnProcessor.forward literally
// (animatediff/models/attention_processor.py:573-619): natural-log softmax over the keys of
// all partners, bias from the row table, probabilities kept in fp32.  One warp per
// (row, head, query).  Test infrastructure on the device side -- never on the product path.
#include "common.cuh"

namespace cvd {

constexpr int kSimtWarps = 4;

template <typename OutT>
__global__ void __launch_bounds__(kSimtWarps * 32)
epi_attn_simt_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ vt, const float* __restrict__ rowp,
                     const int* __restrict__ partner, OutT* __restrict__ out, int R, int N, int heads, int d,
                     int n_partners, int feat, float S) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * d, DV = (d / 16 + 1) * 16, M = n_partners * N;
  float* qs = smem + warp * (160 + M);   // q row (fp32), then the M scores
  float* sc = qs + 160;
  const long long item = (long long)blockIdx.x * kSimtWarps + warp;
  if (item >= (long long)R * heads * N) return;
  const int n = (int)(item % N);
  const int h = (int)((item / N) % heads);
  const int r = (int)(item / ((long long)N * heads));
  for (int i = lane; i < d; i += 32) qs[i] = __bfloat162float(q[((size_t)r * N + n) * C + h * d + i]);
  __syncwarp();
  const float scale = rsqrtf((float)d);
  const float step = S / (float)feat;
  const float xq = pixel_centre(n % feat, step), yq = pixel_centre(n / feat, step);
  float mx = -INFINITY;
  for (int p = 0; p < n_partners; ++p) {
    const int rr = partner[p * R + r];
    float a = 0.f, b = 0.f, c = 0.f, nrm = 1.f, band = 0.f, w = 0.f;
    if (rowp) {
      const float* P = rowp + (size_t)(p * R + r) * CVD_ROWP_STRIDE;
      a = fmaf(P[1], yq, P[0] * xq) + P[2];
      b = fmaf(P[4], yq, P[3] * xq) + P[5];
      c = fmaf(P[7], yq, P[6] * xq) + P[8];
      nrm = sqrtf(a * a + b * b) + 1e-6f;
      band = P[9];
      w = P[10];
    }
    for (int kk = lane; kk < N; kk += 32) {
      const __nv_bfloat16* kr = k + ((size_t)rr * N + kk) * C + h * d;
      float dot = 0.f;
      for (int i = 0; i < d; ++i) dot = fmaf(qs[i], __bfloat162float(kr[i]), dot);
      float bias = 0.f;
      if (rowp) {
        const float dist = fabsf(fmaf(b, pixel_centre(kk / feat, step), a * pixel_centre(kk % feat, step)) + c) / nrm;
        bias = -fmaxf(dist - band, 0.f) * w;
        if (!(fabsf(bias) <= 3.0e38f)) bias = 0.f;
      }
      const float s = dot * scale + bias;
      sc[p * N + kk] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < M; i += 32) {
    const float e = expf(sc[i] - mx);
    sc[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int i = 0; i < d; ++i) {
    float acc = 0.f;
    for (int p = 0; p < n_partners; ++p) {
      const int rr = partner[p * R + r];
      const __nv_bfloat16* vr = vt + (((size_t)rr * heads + h) * DV + i) * N;
      for (int kk = lane; kk < N; kk += 32) acc = fmaf(sc[p * N + kk], __bfloat162float(vr[kk]), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) store_from_float(out, ((size_t)r * N + n) * C + h * d + i, acc * inv);
  }
}

}  // namespace cvd

extern "C" int cvd_epi_attn_fwd_simt(const void* q, const void* k, const void* vt, const float* rowp,
                                     const int* partner, void* out, int out_dtype, int R, int R_kv, int N, int heads,
                                     int head_dim, int n_partners, int feat, float F_mat_size,
                                     cvd_stream_t stream) {
  (void)R_kv;   // rows of k/vt: only bounds the partner indices, which the caller guarantees
  using namespace cvd;
  CVD_REQUIRE(q && k && vt && partner && out, "cvd_epi_attn_fwd_simt: null pointer");
  CVD_REQUIRE(R > 0 && heads > 0 && head_dim > 0 && head_dim <= 160 && n_partners >= 1,
              "cvd_epi_attn_fwd_simt: R=%d heads=%d head_dim=%d n_partners=%d", R, heads, head_dim, n_partners);
  CVD_REQUIRE(feat > 0 && N == feat * feat, "cvd_epi_attn_fwd_simt: N=%d is not feat^2 (feat=%d)", N, feat);
  const size_t smem = (size_t)kSimtWarps * (160 + (size_t)n_partners * N) * sizeof(float);
  CVD_REQUIRE(smem <= 200 * 1024, "cvd_epi_attn_fwd_simt: %d partners x %d keys do not fit shared memory", n_partners, N);
  const long long items = (long long)R * heads * N;
  const unsigned grid = (unsigned)((items + kSimtWarps - 1) / kSimtWarps);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == CVD_DTYPE_F32) {
    auto kern = epi_attn_simt_kernel<float>;
    CVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kSimtWarps * 32, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                              (const __nv_bfloat16*)vt, rowp, partner, (float*)out, R, N, heads,
                                              head_dim, n_partners, feat, F_mat_size);
  } else if (out_dtype == CVD_DTYPE_BF16) {
    auto kern = epi_attn_simt_kernel<__nv_bfloat16>;
    CVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kSimtWarps * 32, smem, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                              (const __nv_bfloat16*)vt, rowp, partner, (__nv_bfloat16*)out, R, N,
                                              heads, head_dim, n_partners, feat, F_mat_size);
  } else {
    return fail(CVD_E_INVALID, "cvd_epi_attn_fwd_simt: out_dtype %d not supported", out_dtype);
  }
  CVD_LAUNCH_CHECK("epi_attn_simt_kernel");
  return CVD_OK;
}
