// Epipolar bias geometry: EpiEncoding.get_attn_map of the reference
// (animatediff/models/epi_module.py:293-346) reduced to a 12-float table per row, plus a
// dense writer that exists only so the bias can be compared element-wise with the reference.
#include "common.cuh"

namespace cvd {

__device__ __forceinline__ void line_of_query(const float* M, float x, float y, float& a, float& b, float& c) {
  // l_q = M (x, y, 1)^T   (torch.bmm(F_mats, coords^T), epi_module.py:315)
  a = fmaf(M[1], y, M[0] * x) + M[2];
  b = fmaf(M[4], y, M[3] * x) + M[5];
  c = fmaf(M[7], y, M[6] * x) + M[8];
}

// One CTA per row.  max_{q,k} d[r,q,k] needs no N^2 pass: |a x + b y + c| is convex in (x,y),
// so for every query the maximum over the key grid sits on one of the 4 grid corners.
__global__ void __launch_bounds__(128) epi_row_params_kernel(const float* __restrict__ F_mats,
                                                             const float* __restrict__ theta,
                                                             float* __restrict__ rowp, int feat, float S,
                                                             float pixel_band, float decay_alpha,
                                                             int first_frame_rule, int frames_per_video) {
  const int r = blockIdx.x;
  __shared__ float M[9];
  __shared__ float red[4];
  __shared__ int bad[4];
  pdl_trigger();      // the attention kernel behind this one may take free slots and run its prologue
  pdl_wait();         // F_mats / theta may come from the kernel in front; rowp may still be read by it
  if (threadIdx.x < 9) {
    float v = F_mats[(size_t)r * 9 + threadIdx.x];
    if (first_frame_rule != CVD_FF_NONE && (r % frames_per_video) == 0) {
      // pseudo line through the query pixel, written as a matrix acting on (x,y,1):
      //   horizontal : (0, -1, y)                       (epi_module.py:341-345)
      //   slope t    : (cos t, sin t, -(cos t x + sin t y))   (epi_module.py:335-339)
      float ct = 0.f, st = -1.f;
      if (first_frame_rule == CVD_FF_RANDOM_SLOPE) {
        float t = theta[0];
        ct = cosf(t);
        st = sinf(t);
      }
      const float m[9] = {0.f, 0.f, ct, 0.f, 0.f, st, -ct, -st, 0.f};
      v = m[threadIdx.x];
    }
    M[threadIdx.x] = v;
  }
  __syncthreads();
  const float step = S / (float)feat;
  const float lo = pixel_centre(0, step), hi = pixel_centre(feat - 1, step);
  float dmax = 0.f;
  int nonfinite = 0;
  for (int q = threadIdx.x; q < feat * feat; q += blockDim.x) {
    const float x = pixel_centre(q % feat, step), y = pixel_centre(q / feat, step);
    float a, b, c;
    line_of_query(M, x, y, a, b, c);
    const float nrm = sqrtf(a * a + b * b) + 1e-6f;
    float t = fabsf(fmaf(b, lo, a * lo) + c);
    t = fmaxf(t, fabsf(fmaf(b, lo, a * hi) + c));
    t = fmaxf(t, fabsf(fmaf(b, hi, a * lo) + c));
    t = fmaxf(t, fabsf(fmaf(b, hi, a * hi) + c));
    const float d = t / nrm;
    // torch.max propagates NaN and inf-inf turns the whole row into NaN, which the reference
    // then scrubs to 0 (epi_module.py:404-407): a non-finite row carries no bias at all.
    if (!(fabsf(d) <= 3.0e38f) || !(fabsf(a) <= 3.0e38f) || !(fabsf(b) <= 3.0e38f)) nonfinite = 1;
    dmax = fmaxf(dmax, d);
  }
  dmax = warp_max(dmax);
  nonfinite = __any_sync(0xffffffffu, nonfinite);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = dmax;
    bad[threadIdx.x >> 5] = nonfinite;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    dmax = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    nonfinite = bad[0] | bad[1] | bad[2] | bad[3];
    float* o = rowp + (size_t)r * CVD_ROWP_STRIDE;
    if (nonfinite) {
      for (int i = 0; i < CVD_ROWP_STRIDE; ++i) o[i] = 0.f;
    } else {
      const float band = pixel_band / floorf(S * 0.5f) * dmax;   // pixel_band / (S // 2) * max
      for (int i = 0; i < 9; ++i) o[i] = M[i];
      o[9] = band;
      o[10] = decay_alpha / (band + 1e-6f);
      o[11] = 0.f;
    }
  }
}

// bias[r,q,k] = -max(d - band, 0) * w     (epi_module.py:322-327); one CTA per (row, query).
__global__ void __launch_bounds__(128) epi_bias_dense_kernel(const float* __restrict__ rowp,
                                                             float* __restrict__ bias, int feat, float S) {
  const int N = feat * feat;
  const int r = blockIdx.y, q = blockIdx.x;
  const float* p = rowp + (size_t)r * CVD_ROWP_STRIDE;
  const float step = S / (float)feat;
  float a, b, c;
  line_of_query(p, pixel_centre(q % feat, step), pixel_centre(q / feat, step), a, b, c);
  const float nrm = sqrtf(a * a + b * b) + 1e-6f;
  const float band = p[9], w = p[10];
  float* out = bias + ((size_t)r * N + q) * N;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const float d = fabsf(fmaf(b, pixel_centre(k / feat, step), a * pixel_centre(k % feat, step)) + c) / nrm;
    float v = -fmaxf(d - band, 0.f) * w;
    if (!(fabsf(v) <= 3.0e38f)) v = 0.f;
    out[k] = v;
  }
}

}  // namespace cvd

extern "C" int cvd_epi_row_params(const float* F_mats, const float* theta, float* rowp, int rows, int feat,
                                  float F_mat_size, float pixel_band, float decay_alpha, int first_frame_rule,
                                  int frames_per_video, cvd_stream_t stream) {
  CVD_REQUIRE(F_mats && rowp, "cvd_epi_row_params: null pointer");
  CVD_REQUIRE(rows > 0 && feat > 0 && feat <= 64, "cvd_epi_row_params: rows=%d feat=%d out of range", rows, feat);
  CVD_REQUIRE(first_frame_rule >= CVD_FF_NONE && first_frame_rule <= CVD_FF_RANDOM_SLOPE,
              "cvd_epi_row_params: unknown first_frame_rule %d", first_frame_rule);
  CVD_REQUIRE(first_frame_rule != CVD_FF_RANDOM_SLOPE || theta, "cvd_epi_row_params: theta required");
  CVD_REQUIRE(frames_per_video > 0, "cvd_epi_row_params: frames_per_video must be positive");
  CVD_CUDA(cvd::launch_pdl(cvd::epi_row_params_kernel, dim3(rows), dim3(128), 0, (cudaStream_t)stream, F_mats, theta, rowp,
                           feat, F_mat_size, pixel_band, decay_alpha, first_frame_rule, frames_per_video));
  CVD_LAUNCH_CHECK("epi_row_params_kernel");
  return CVD_OK;
}

extern "C" int cvd_epi_bias_dense(const float* rowp, float* bias, int rows, int feat, float F_mat_size,
                                  cvd_stream_t stream) {
  CVD_REQUIRE(rowp && bias, "cvd_epi_bias_dense: null pointer");
  CVD_REQUIRE(rows > 0 && rows <= 65535 && feat > 0 && feat <= 64, "cvd_epi_bias_dense: rows=%d feat=%d", rows, feat);
  dim3 grid(feat * feat, rows);
  cvd::epi_bias_dense_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(rowp, bias, feat, F_mat_size);
  CVD_LAUNCH_CHECK("epi_bias_dense_kernel");
  return CVD_OK;
}
