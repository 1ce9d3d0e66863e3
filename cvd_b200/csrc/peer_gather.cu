// Partner-row exchange for the view-sharded epi attention: every rank PULLS the rows its queries
// attend to straight out of its peers' HBM over NVLink (peer pointers from symmetric memory), with
// 16-byte loads, instead of all-gathering every rank's rows through NCCL.  Only the rows that are
// needed cross the links (R_local rows per call whatever the world size), there is no collective
// launch latency, and the data lands in one contiguous local buffer that the attention kernel's TMA
// reads from local HBM / L2 (eight CTAs share every K/V tile: reading them from the peer inside the
// attention kernel would cross NVLink eight times, peer reads bypass the local L2).
#include "common.cuh"

namespace cvd {

struct PeerBases {
  const uint4* base[8];
};

// grid: (chunks per row, rows).  dst[r] = src[partner[r] / rows_per_src][partner[r] % rows_per_src]
__global__ void __launch_bounds__(256)
peer_gather_rows_kernel(PeerBases src, const int* __restrict__ partner, uint4* __restrict__ dst, int rows_per_src,
                        size_t row_vec /* uint4 per row */) {
  const int r = blockIdx.y;
  const int g = partner[r];
  const uint4* s = src.base[g / rows_per_src] + (size_t)(g % rows_per_src) * row_vec;
  uint4* d = dst + (size_t)r * row_vec;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < row_vec; i += stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k < row_vec) v[k] = s[i + k];     // 4 x 16-byte loads in flight per thread (NVLink latency ~2 us)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k < row_vec) d[i + k] = v[k];
  }
}

}  // namespace cvd

extern "C" int cvd_peer_gather_rows(const void* const* src_bases, int n_src, const int* partner, void* dst, int rows,
                                    int rows_per_src, long long row_bytes, cvd_stream_t stream) {
  using namespace cvd;
  CVD_REQUIRE(src_bases && partner && dst, "cvd_peer_gather_rows: null pointer");
  CVD_REQUIRE(n_src >= 1 && n_src <= 8, "cvd_peer_gather_rows: n_src=%d (1..8 ranks of one NVSwitch box)", n_src);
  CVD_REQUIRE(rows > 0 && rows <= 65535 && rows_per_src > 0 && row_bytes > 0 && row_bytes % 16 == 0,
              "cvd_peer_gather_rows: rows=%d rows_per_src=%d row_bytes=%lld (must be a multiple of 16)", rows, rows_per_src,
              row_bytes);
  PeerBases pb;
  for (int i = 0; i < 8; ++i) pb.base[i] = (const uint4*)src_bases[i < n_src ? i : 0];
  for (int i = 0; i < n_src; ++i)
    CVD_REQUIRE(src_bases[i] && (((uintptr_t)src_bases[i]) & 15) == 0, "cvd_peer_gather_rows: base %d is null or unaligned", i);
  const size_t row_vec = (size_t)row_bytes / 16;
  // enough CTAs to keep ~all SMs pulling: rows x chunks ~ 4 waves
  int chunks = (int)((row_vec + 256 * 4 * 4 - 1) / (256 * 4 * 4));
  const int want = (4 * kNumSMs + rows - 1) / rows;
  if (chunks > want) chunks = want;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, rows);
  peer_gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pb, partner, (uint4*)dst, rows_per_src, row_vec);
  CVD_LAUNCH_CHECK("peer_gather_rows_kernel");
  return CVD_OK;
}
