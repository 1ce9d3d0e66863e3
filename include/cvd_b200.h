/*
 * cvd_b200 -- C ABI of the B200-native CVD cross-video synchronisation hot path.
 *
 * The reference (JimmyPang02/CVD) is pure Python/PyTorch and has no FFI of its own; the
 * entry points below are what a binding for this path would call, one per stage of the
 * reference's EpiModule / VanillaTemporalModule forward.  Each declaration cites the
 * reference lines it replaces (paths relative to the reference checkout).  The ctypes
 * binding the Python host side uses is cvd_b200/_lib.py; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates inputs,
 *     outputs and workspaces); the library never allocates or frees caller memory and
 *     keeps no state besides per-process kernel attributes;
 *   - all tensors are dense row-major; "bf16" buffers are passed as const void*;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); every call is
 *     asynchronous and stream-ordered, no call synchronises the device;
 *   - return value 0 = success, negative = error (CVD_E_*); cvd_last_error() returns a
 *     thread-local message.  Unsupported shapes are errors -- there is no fallback path.
 */
#ifndef CVD_B200_H_
#define CVD_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CVD_OK 0
#define CVD_E_INVALID (-1)   /* bad argument / unsupported shape */
#define CVD_E_CUDA (-2)      /* CUDA runtime or driver error      */

#define CVD_DTYPE_F32 0
#define CVD_DTYPE_BF16 1
#define CVD_DTYPE_F16 2

/* token layouts of a [B,C,F,h,w] video tensor */
#define CVD_LAYOUT_BF_HW 0   /* "(b f) (h w) c"  -- epi module,      epi_module.py:165    */
#define CVD_LAYOUT_BHW_F 1   /* "(b h w) f c"    -- temporal module, motion_module.py:378 */

/* first-frame rule of the epipolar bias (epi_module.py:316,331-346) */
#define CVD_FF_NONE 0        /* use F_r on every row (not reachable in the reference)   */
#define CVD_FF_HORIZONTAL 1  /* rows r % frames == 0: line (0,-1,y_q)                    */
#define CVD_FF_RANDOM_SLOPE 2/* rows r % frames == 0: line through q with slope *theta   */

#define CVD_ROWP_STRIDE 12   /* floats per row of the row-parameter table               */

typedef void* cvd_stream_t;

const char* cvd_last_error(void);
int cvd_abi_version(void);

/* ---- a1/a2: epipolar bias, EpiEncoding.get_attn_map (epi_module.py:293-329) + NaN scrub
 *      (epi_module.py:404-407).  The dense [rows,N,N] tensor is never built on the product
 *      path: this prepass reduces each row's geometry to 12 floats
 *        rowp[r] = { M00..M22 (effective line matrix, l_q = M x_q), band_r, w_r, 0 }
 *      band_r = pixel_band/(S//2) * max_{q,k} d,  w_r = decay_alpha/(band_r+1e-6).
 *      `theta` is a device scalar (radians) used when first_frame_rule == RANDOM_SLOPE. */
int cvd_epi_row_params(const float* F_mats, const float* theta, float* rowp, int rows, int feat,
                       float F_mat_size, float pixel_band, float decay_alpha,
                       int first_frame_rule, int frames_per_video, cvd_stream_t stream);

/* Parity/debug only: materialise bias[rows, N, N] (fp32) from the row parameters. */
int cvd_epi_bias_dense(const float* rowp, float* bias, int rows, int feat, float F_mat_size,
                       cvd_stream_t stream);

/* ---- a4/a5: EpiAttnProcessor core (attention_processor.py:573-619,637):
 *        out[r,n,h*d:(h+1)*d] = softmax_k( q.k/sqrt(d) + bias[r,n,k] ) v
 *      with K/V taken from row partner[p*R + r] and bias row p*R + r for partner
 *      p < n_partners (keys of all partners form one softmax, attention_processor.py:579-583).
 *        q    : bf16 [R, N, heads*head_dim]
 *        k    : bf16 [R_kv, N, heads*head_dim]   R_kv >= R when K/V were gathered from other
 *               ranks (view sharding): partner[] then holds rows of the gathered buffer
 *        vt   : bf16 [R_kv, heads, DV, N], DV = (head_dim/16 + 1)*16: V transposed per head; row
 *               head_dim must be 1.0 and the rows above it 0 (the kernel reads the softmax
 *               denominator from that row)
 *        rowp : fp32 [n_partners*R, 12] from cvd_epi_row_params, or NULL for no bias
 *               (epi_no_attention_mask, epi_module.py:392)
 *        partner : int32 [n_partners*R]
 *        out  : bf16 [R, N, heads*head_dim]
 *      tcgen05/TMEM/TMA kernel; head_dim in {40, 80, 160}, N = feat*feat, feat in {4,8,16,32,64}. */
int cvd_epi_attn_fwd(const void* q, const void* k, const void* vt, const float* rowp,
                     const int* partner, void* out, int R, int R_kv, int N, int heads, int head_dim,
                     int n_partners, int feat, float F_mat_size, cvd_stream_t stream);

/* ---- (e) view-sharded attention: K / V^T spread over several buffers, partner rows numbered globally.
 *      Replaces "all-gather K / V of the partner views, then attend" for attention_processor.py:573-583 when the views
 *      of a V-view run are sharded over the GPUs of one NVSwitch box (the reference has no multi-GPU inference).
 *      k_srcs / vt_srcs: n_src (1..8) device pointers, each a [rows_per_src, N, heads*head_dim] /
 *      [rows_per_src, heads, DV, N] buffer -- e.g. the symmetric-memory copies of every rank through their peer (NVLink)
 *      mappings, own rank included; `partner` holds GLOBAL rows: row p = source p / rows_per_src, local row
 *      p % rows_per_src.  Partner tiles are read straight out of the owning buffer by TMA (unicast: measured safe over
 *      NVLink; use it for latents <= 8x8, where a (row, head) has ONE query tile and every byte crosses once).  When all
 *      sources are memory of the calling device and a (row, head) has several query tiles, those CTAs form a
 *      thread-block cluster and each K / V^T tile is fetched once and multicast into the shared memory of all of them
 *      (never for peer memory: multicast TMA from a peer mapping wedges the GPU on driver 580.159).
 *      The caller orders the call after the peers' writes (one cross-rank barrier).  Everything else as
 *      cvd_epi_attn_fwd. */
int cvd_epi_attn_fwd_sharded(const void* q, const void* const* k_srcs, const void* const* vt_srcs, int n_src,
                             int rows_per_src, const float* rowp, const int* partner, void* out, int R, int N,
                             int heads, int head_dim, int n_partners, int feat, float F_mat_size,
                             cvd_stream_t stream);

/* Bring-up aid: as cvd_epi_attn_fwd, and additionally dumps the raw Q K^T accumulator of key
 * tile 0 of every CTA into dbg_s (fp32 [grid, 128, 64], grid = R*heads*ceil(N/128)), followed by
 * int64 [grid, 32, 8] clock64 stamps of thread 0 at the hand-off points of each key tile. */
int cvd_epi_attn_fwd_debug(const void* q, const void* k, const void* vt, const float* rowp,
                           const int* partner, void* out, float* dbg_s, int R, int R_kv, int N,
                           int heads, int head_dim, int n_partners, int feat, float F_mat_size,
                           cvd_stream_t stream);

/* ---- f3 (first CUDA path): forward with the statistic the backward needs, and the backward of the attention core.
 *      Replaces autograd through EpiAttnProcessor (attention_processor.py:573-619) in train_epi_control.py:512-642; the
 *      epipolar bias has no gradient (epi_module.py:329,393).
 *      cvd_epi_attn_fwd_lse: as cvd_epi_attn_fwd, plus lse fp32 [R, heads, N] = log2 sum_k 2^(s_k log2e/sqrt(d) + bias_k log2e).
 *      cvd_epi_attn_bwd: q [R,N,C], k / v [R_kv,N,C] bf16 in their NATURAL layout, out / d_out [R,N,C] bf16, rowp fp32
 *      [R,12] or NULL, partner int32 [R] (single partner), lse from the forward; delta: fp32 workspace [R, heads, N];
 *      dq [R,N,C], dk / dv [R_kv,N,C] fp32 are ACCUMULATED (zero-fill them first).  head_dim in {40, 80, 160},
 *      N a multiple of 16. */
int cvd_epi_attn_fwd_lse(const void* q, const void* k, const void* vt, const float* rowp,
                         const int* partner, void* out, float* lse, int R, int R_kv, int N, int heads,
                         int head_dim, int n_partners, int feat, float F_mat_size, cvd_stream_t stream);
int cvd_epi_attn_bwd(const void* q, const void* k, const void* v, const float* rowp, const int* partner,
                     const void* out, const void* d_out, const float* lse, float* delta, float* dq,
                     float* dk, float* dv, int R, int R_kv, int N, int heads, int head_dim, int feat,
                     float F_mat_size, cvd_stream_t stream);

/* Same contract (out may be bf16 or fp32), plain CUDA-core fp32 kernel that follows the
 * reference's arithmetic literally (natural-log softmax, unrounded probabilities): the
 * on-device cross-check used by the tests at sizes where the CPU oracle is slow.  Any
 * head_dim <= 160.  Never called by the product path. */
int cvd_epi_attn_fwd_simt(const void* q, const void* k, const void* vt, const float* rowp,
                          const int* partner, void* out, int out_dtype, int R, int R_kv, int N,
                          int heads, int head_dim, int n_partners, int feat, float F_mat_size,
                          cvd_stream_t stream);

/* ---- a7/a10: InflatedGroupNorm + rearrange to tokens (resnet.py:30-40, epi_module.py:163-165,
 *      motion_module.py:377-378).  x [B,C,F,HW] (x_dtype) -> y bf16 tokens in `layout`.
 *      stats: fp32 workspace [B*F*groups*2]. */
int cvd_groupnorm_tokens_fwd(const void* x, int x_dtype, const float* gamma, const float* beta,
                             void* y, float* stats, int B, int C, int F, int HW, int groups,
                             float eps, int layout, cvd_stream_t stream);

/* ---- a6: LayerNorm over C of fp32 tokens -> bf16 (epi_module.py:254,266; motion_module.py:449,457).
 *      Optional pe[seq_len, C] is added after the norm (PositionalEncoding, motion_module.py:479-481)
 *      with position = token % seq_len. */
int cvd_layernorm_fwd(const float* x, const float* gamma, const float* beta, const float* pe,
                      void* y, int tokens, int C, int seq_len, float eps, cvd_stream_t stream);

/* ---- a6: fused residual update + LayerNorm: x += (y + bias) * scale in place (fp32; y bf16 or NULL, bias
 *      fp32[C] or NULL), then n_out = LayerNorm(x) (+pe) as bf16.  Replaces the `x = attn(norm(x)) + x` /
 *      `norm(x)` pairs of EpiTransformerBlock / TemporalTransformerBlock (epi_module.py:254-268,
 *      motion_module.py:449-459) with one pass over the residual stream. */
int cvd_residual_layernorm_fwd(float* x, const void* y, const float* bias, float scale,
                               const float* gamma, const float* beta, const float* pe, void* n_out,
                               int tokens, int C, int seq_len, float eps, cvd_stream_t stream);

/* ---- a6: GEGLU gate (diffusers FeedForward used at epi_module.py:248):
 *      out[t, i] = hg[t, i] * gelu_erf(hg[t, inner + i]),  hg bf16 [T, 2*inner] -> out bf16 [T, inner] */
int cvd_geglu_fwd(const void* hg, void* out, int tokens, int inner, cvd_stream_t stream);

/* ---- residual stream update: x[t,c] += (y[t,c] + bias[c]) * scale, x fp32 in place, y bf16;
 *      optional bf16 copy of the result (input of the next GEMM).  (epi_module.py:263,266;
 *      attention_processor.py:641-648: bias = to_out bias, scale = 1/rescale_output_factor) */
int cvd_residual_add_fwd(float* x, const void* y, const float* bias, float scale, void* x_bf16,
                         int tokens, int C, cvd_stream_t stream);

/* ---- a7: proj_out epilogue: tokens -> "b c f h w" + residual (epi_module.py:185-187,
 *      motion_module.py:390-392).  out = residual + transpose(tok + bias). */
int cvd_tokens_to_video_fwd(const void* tok, const float* bias, const void* residual, void* out,
                            int io_dtype, int B, int C, int F, int HW, int layout,
                            cvd_stream_t stream);

/* ---- a9: temporal self-attention over the F frames of every pixel
 *      (attention_processor.py:84-94 / :346-357).  q,k,v,out bf16 [T, F, heads*head_dim], F <= 32. */
int cvd_temporal_attn_fwd(const void* q, const void* k, const void* v, void* out, int T, int F,
                          int heads, int head_dim, cvd_stream_t stream);

/* ---- a9: CameraCtrl pose merge epilogue (attention_processor.py:324-326):
 *      out = merged * scale + x,  merged = qkv_merge(x + pose) computed by the caller's GEMM.
 *      x_plus_pose helper: out_bf16 = x_bf16 + pose (pose in video layout [B,C,F,HW], x tokens "(b h w) f c"). */
int cvd_add_pose_tokens_fwd(const void* x_tok, const void* pose, int pose_dtype, void* out_tok,
                            int B, int C, int F, int HW, cvd_stream_t stream);
int cvd_scale_add_fwd(const void* a, const void* b, float scale, void* out, long long n,
                      cvd_stream_t stream);

/* ---- a11 on the device: per-row fundamental matrices of a view pairing, replacing the Python loop the
 *      V-view sampler runs on the CPU before every UNet call
 *      (pipelines/pipeline_animation_epi_advanced.py:636-647 -> data/dataset_validation.py:28-64):
 *        F[s] = K[t]^-T E(inv(c2w[t]) c2w[s]) K[s]^-1,  t = partner[s]
 *      c2w fp64 [rows,4,4] (any affine pose: bottom row 0 0 0 1; inverted like np.linalg.inv), K fp64 [rows,3,3], partner int32 [rows], F_mats fp32 [rows,3,3]. */
int cvd_pairwise_fundamental(const double* c2w, const double* K, const int* partner, float* F_mats,
                             int rows, cvd_stream_t stream);

/* ---- f2, second half: Pluecker ray embedding of the pose encoder input, replacing ray_condition
 *      (data/dataset_validation.py:102-144; data/dataset_train_realestate10k.py:168-210) on the CPU.
 *      intrinsics fp32 [views,4] = (fx, fy, cx, cy) in pixels of the H x W grid; c2w fp32 [views,4,4];
 *      flip_flag uint8 [views] or NULL (horizontally mirrored pixel grid, :117-126);
 *      out fp32 [views,6,H,W] (channel_major != 0: what the dataset passes on after .permute(0,3,1,2)) or
 *      [views,H,W,6] (ray_condition's own layout): (o x d, d) with d the world-space unit ray through the pixel centre. */
int cvd_plucker_embedding(const float* intrinsics, const float* c2w, const unsigned char* flip_flag, float* out,
                          int views, int H, int W, int channel_major, cvd_stream_t stream);

/* ---- multi-GPU view sharding (new; the reference has no multi-GPU inference): pull the rows this
 *      rank's queries attend to out of the peers' memory over NVLink.
 *        dst[r] = src_bases[partner[r] / rows_per_src][partner[r] % rows_per_src],  r < rows
 *      src_bases: HOST array of n_src (<= 8) device pointers (peer-mapped, e.g. symmetric memory);
 *      partner: device int32 [rows] of GLOBAL row numbers; rows are row_bytes (multiple of 16) long.
 *      The caller orders it after the peers' writes (cross-rank barrier). */
int cvd_peer_gather_rows(const void* const* src_bases, int n_src, const int* partner, void* dst,
                         int rows, int rows_per_src, long long row_bytes, cvd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CVD_B200_H_ */
