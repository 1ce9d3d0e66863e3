// Pose -> fundamental matrices on the device: the per-UNet-call loop of the V-view sampler
// (animatediff/pipelines/pipeline_animation_epi_advanced.py:636-647 calling
// animatediff/data/dataset_validation.py:28-64) as one kernel launch.  One thread per row:
//   s2t = inv(c2w[tgt]) c2w[src];  E = R [(-R^T t)]x;  F = K_tgt^-T E K_src^-1      (fp64, like numpy)
#include "common.cuh"

namespace cvd {

__device__ void inv3(const double* m, double* o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double r = 1.0 / det;
  o[0] = A * r; o[1] = -(b * i - c * h) * r; o[2] = (b * f - c * e) * r;
  o[3] = B * r; o[4] = (a * i - c * g) * r;  o[5] = -(a * f - c * d) * r;
  o[6] = C * r; o[7] = -(a * h - b * g) * r; o[8] = (a * e - b * d) * r;
}
__device__ void mul3(const double* a, const double* b, double* o) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[r * 3 + c] = a[r * 3] * b[c] + a[r * 3 + 1] * b[3 + c] + a[r * 3 + 2] * b[6 + c];
}

__global__ void pairwise_fundamental_kernel(const double* __restrict__ c2w, const double* __restrict__ K,
                                            const int* __restrict__ partner, float* __restrict__ F, int rows) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= rows) return;
  const int t = partner[s];
  const double* Ms = c2w + (size_t)s * 16;
  const double* Mt = c2w + (size_t)t * 16;
  // inverse of the target pose as an affine map: inv([A|p]) = [A^-1 | -A^-1 p] -- what np.linalg.inv returns for any
  // 4x4 with bottom row (0,0,0,1), orthonormal A or not (scaled / sheared user poses included)
  double At[9], Ati[9], Rs[9], R[9], pt[3], ps[3], tr[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { At[r * 3 + c] = Mt[r * 4 + c]; Rs[r * 3 + c] = Ms[r * 4 + c]; }
    pt[r] = Mt[r * 4 + 3];
    ps[r] = Ms[r * 4 + 3];
  }
  inv3(At, Ati);
  mul3(Ati, Rs, R);                                  // linear part of s2t
  for (int r = 0; r < 3; ++r)                        // translation of s2t = A_t^-1 (p_s - p_t)
    tr[r] = Ati[r * 3] * (ps[0] - pt[0]) + Ati[r * 3 + 1] * (ps[1] - pt[1]) + Ati[r * 3 + 2] * (ps[2] - pt[2]);
  double te[3];                                      // t_ess = -R^T t
  for (int r = 0; r < 3; ++r) te[r] = -(R[r] * tr[0] + R[3 + r] * tr[1] + R[6 + r] * tr[2]);
  const double X[9] = {0.0, -te[2], te[1], te[2], 0.0, -te[0], -te[1], te[0], 0.0};
  double E[9], K1i[9], K2i[9], K2iT[9], T1[9], Fm[9];
  mul3(R, X, E);
  inv3(K + (size_t)s * 9, K1i);
  inv3(K + (size_t)t * 9, K2i);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) K2iT[r * 3 + c] = K2i[c * 3 + r];
  mul3(K2iT, E, T1);
  mul3(T1, K1i, Fm);
  for (int i = 0; i < 9; ++i) F[(size_t)s * 9 + i] = (float)Fm[i];
}

// Pluecker ray embedding of the CameraCtrl pose encoder input: ray_condition of the reference
// (animatediff/data/dataset_validation.py:102-144) for V cameras at H x W, written channel-major [V, 6, H, W] (the layout
// the dataset hands on after its .permute(0, 3, 1, 2), :256-257) or pixel-major [V, H, W, 6] (ray_condition's own).
// One thread per (camera, pixel); fp32 like the reference (its c2w / intrinsics are float tensors).
//   d = normalize((i + .5 - cx) / fx, (j + .5 - cy) / fy, 1);  rays_d = R d;  o = c2w[:3, 3];  out = (o x rays_d, rays_d)
__global__ void plucker_kernel(const float* __restrict__ K4, const float* __restrict__ c2w, const unsigned char* __restrict__ flip,
                               float* __restrict__ out, int V, int H, int W, int channel_major) {
  const int HW = H * W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)V * HW) return;
  const int v = (int)(idx / HW), pix = (int)(idx % HW);
  const int jj = pix / W;
  int ii = pix % W;
  if (flip != nullptr && flip[v]) ii = W - 1 - ii;     // horizontally flipped grid (:117-126)
  const float fx = K4[v * 4], fy = K4[v * 4 + 1], cx = K4[v * 4 + 2], cy = K4[v * 4 + 3];
  const float xs = ((float)ii + 0.5f - cx) / fx, ys = ((float)jj + 0.5f - cy) / fy;
  const float nrm = sqrtf(xs * xs + ys * ys + 1.0f);
  const float dx = xs / nrm, dy = ys / nrm, dz = 1.0f / nrm;
  const float* M = c2w + (size_t)v * 16;
  const float rx = M[0] * dx + M[1] * dy + M[2] * dz;
  const float ry = M[4] * dx + M[5] * dy + M[6] * dz;
  const float rz = M[8] * dx + M[9] * dy + M[10] * dz;
  const float ox = M[3], oy = M[7], oz = M[11];
  const float val[6] = {oy * rz - oz * ry, oz * rx - ox * rz, ox * ry - oy * rx, rx, ry, rz};
  if (channel_major) {
#pragma unroll
    for (int c = 0; c < 6; ++c) out[((size_t)v * 6 + c) * HW + pix] = val[c];
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c) out[((size_t)v * HW + pix) * 6 + c] = val[c];
  }
}

}  // namespace cvd

extern "C" int cvd_plucker_embedding(const float* intrinsics, const float* c2w, const unsigned char* flip_flag, float* out,
                                     int views, int H, int W, int channel_major, cvd_stream_t stream) {
  CVD_REQUIRE(intrinsics && c2w && out, "cvd_plucker_embedding: null pointer");
  CVD_REQUIRE(views > 0 && H > 0 && W > 0, "cvd_plucker_embedding: views=%d H=%d W=%d", views, H, W);
  const long long total = (long long)views * H * W;
  cvd::plucker_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(intrinsics, c2w, flip_flag, out, views,
                                                                                       H, W, channel_major);
  CVD_LAUNCH_CHECK("plucker_kernel");
  return CVD_OK;
}

extern "C" int cvd_pairwise_fundamental(const double* c2w, const double* K, const int* partner, float* F_mats, int rows,
                                        cvd_stream_t stream) {
  CVD_REQUIRE(c2w && K && partner && F_mats, "cvd_pairwise_fundamental: null pointer");
  CVD_REQUIRE(rows > 0, "cvd_pairwise_fundamental: rows=%d", rows);
  cvd::pairwise_fundamental_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c2w, K, partner, F_mats, rows);
  CVD_LAUNCH_CHECK("pairwise_fundamental_kernel");
  return CVD_OK;
}
