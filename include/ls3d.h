/* ls3d.h - C ABI of lidarseg3d_b200 (B200 / sm_100a kernels for the MSeg3D / SDSeg3D forward path).
 *
 * Conventions (all entry points):
 *   - plain device pointers + sizes, no torch types; the caller owns every buffer (workspaces are
 *     sized by the paired *_workspace_bytes query) and passes the cudaStream_t as `void* stream`;
 *   - returns 0 on success, a negative LS3D_ERR_* for argument errors, a positive cudaError_t for
 *     launch failures; never throws, never allocates, never synchronises the device;
 *   - citations "replaces <file:line>" point into jialeli1/lidarseg3d (the reference).
 */
#ifndef LS3D_H_
#define LS3D_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LS3D_EPI_LINEAR 0
#define LS3D_EPI_ATTN 1

/* ------------------------------------------------------------------------------------------------
 * Gather-GEMM (tcgen05, error-compensated bf16x3 products by default / fp32 accumulate) with fused epilogue.
 *   out[j, :cout] = epi( sum_{k<koff} in[nbr[k][j], :] . W[k]^T )
 * replaces: spconv `indice_conv` as called by det3d/models/backbones/scn_unet.py:15-20,39-46,89-160
 *           (SubMConv3d / SparseConv3d / SparseInverseConv3d + BatchNorm1d + ReLU + residual,
 *           scn_unet.py:57-67,163-187), and the nn.Linear(+BN/LN/ReLU) layers of
 *           det3d/models/point_heads/point_seg_mseg3d_head.py:35-105, context_module.py:60-117,
 *           211-250,320-376, point_seg_batchloss_head.py:33-53, readers/voxel_encoder.py:128-270.
 * The input row may be the channel concatenation [in0 | in1] (UR_block concat, scn_unet.py:166).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ls3d_gemm_args {
  const float* in0;  /* [rows_in, ld0] fp32, first c0 channels of the input row               */
  const float* in1;  /* optional second source (next c1 channels), NULL if c1 == 0            */
  int32_t ld0, c0, ld1, c1; /* c0, c1, ld0, ld1 multiples of 4                                 */
  const int32_t* nbr; /* [koff][m_out] input row per (offset, output row), -1 = none; NULL =   */
                      /* identity (dense Linear, requires koff == 1)                           */
  int32_t koff, m_out;
  const float* w;     /* ls3d_gemm_pack_bf16x3 image (below)                                     */
  int32_t cin_pad;    /* multiple of 16, >= c0 + c1                                            */
  int32_t n_pad;      /* multiple of 16 in [16, 256]                                           */
  int32_t cout;       /* valid output columns (<= n_pad)                                       */
  int32_t epi;        /* LS3D_EPI_LINEAR | LS3D_EPI_ATTN                                       */
  const float* scale; /* [cout] or NULL (=1)  : y = acc * scale + shift  (folded BatchNorm)    */
  const float* shift; /* [cout] or NULL (=0)                                                   */
  int32_t relu;
  const float* res;   /* residual rows [m_out, ld_res] (res_mode 1: add before ReLU,           */
  int32_t ld_res;     /*   2: add after ReLU; 0: none)                                         */
  int32_t res_mode;
  const float* red0;  /* channel_reduction (scn_unet.py:173-187): out[c] += cat[2c] + cat[2c+1] */
  const float* red1;  /*   where cat = [red0 (red_c ch) | red1 (red_c ch)]; NULL = off          */
  int32_t ld_red0, ld_red1, red_c;
  int32_t n_ln;       /* 0, 1 or 2 chained LayerNorms over the cout columns                    */
  const float* ln_g0; const float* ln_b0;
  const float* ln_g1; const float* ln_b1;
  float ln_eps;
  /* LS3D_EPI_ATTN: q = acc + shift; per head softmax(q K^T * attn_scale) V over the frame's
   * class tokens (context_module.py:320-376). K/V: [n_frames][n_head][n_tok][24] fp32.         */
  const float* attn_k; const float* attn_v;
  const int32_t* frame_off; /* [n_frames] first output row of each frame (device)              */
  int32_t n_frames, n_tok, n_head;
  float attn_scale;
  const float* row_mask; /* optional: rows with row_mask[r*ld_mask] != 1.0f are written as zeros  */
  int32_t ld_mask;       /*   (feature completion, point_seg_mseg3d_head.py:314-334)                */
  float* out;         /* [m_out, ld_out]                                                       */
  int32_t ld_out;
  int32_t round_out;  /* must be 0 (tf32 rounding of stored outputs: retired single-pass engine)                  */
  int32_t debug_skip; /* development only, must be 0: bit0 no A gathers, bit1 no W loads, bit2 no MMA, bit3 no epilogue */
  int32_t precise;    /* must be 2: error-compensated bf16x3 - operands split on chip into bf16 hi + lo,                  */
                      /*    x_hi.W_hi + x_hi.W_lo + x_lo.W_hi, fp32 accumulate (~2^-17 per product, fp32-equivalent);    */
                      /*    `w` = ls3d_gemm_pack_bf16x3 image.  (0 / 1 named the retired TF32 / 3xTF32 engines.)         */
  /* Tile plan of the rulebook `nbr` (ls3d_tile_plan_build; NULL = gather per pair).  With a plan, sparse launches run    */
  /* on the gather-once engine (csrc/gather_gemm_once.cu): distinct input rows of a tile staged once, all offsets served  */
  /* from shared memory.  The plan is a pure function of `nbr` and is cached with it (spconv's indice_key cache).         */
  const int32_t* plan_hdr;
  const uint16_t* plan_local;
  const int32_t* plan_pool;
} ls3d_gemm_args;

int ls3d_gather_gemm(const ls3d_gemm_args* args, void* stream);

/* Weight image of the default engine (precise = 2): from fp32 w_kio [koff][cin][cout] on the device (the spconv weight
 * [kz, ky, kx, Cin, Cout] with the offsets flattened; an nn.Linear weight is its transpose with koff = 1) to one block of
 * n_pad * 128 bytes per (offset, 32-channel chunk), the swizzled shared-memory image of a tcgen05 K-major B tile:
 *   n_pad <= 96: 2 n_pad rows of 64 bytes, rows [0, n_pad) = bf16 hi of W[n, 32c .. 32c+31], rows [n_pad, 2 n_pad) = bf16 lo,
 *                16-byte unit u of row r holds logical unit u ^ ((r >> 1) & 3)   (SWIZZLE_64B);
 *   n_pad  > 96: n_pad rows of 128 bytes [hi (32 ch) | lo (32 ch)], unit u of row r holds logical unit u ^ (r & 7);
 *   hi = bf16_rn(w), lo = bf16_rn(w - hi); cin_pad = cin rounded up to 16, n_pad = cout rounded up to 16 (returned).
 * The same chunks are what ls3d_sffm_decoder streams. */
int ls3d_gemm_packed_bytes(int32_t koff, int32_t cin, int32_t cout, int64_t* bytes, int32_t* cin_pad, int32_t* n_pad);
int ls3d_gemm_pack_bf16x3(const float* w_kio, int32_t koff, int32_t cin, int32_t cout, void* out, void* stream);

/* Tile plan of a rulebook table nbr[koff][m_out] (see ls3d_gemm_args.plan_*): per 128-row output tile the sorted list of
 * distinct input rows (in `pool`), the uint16 position table local[tile][k][128] and a 32-int header {n_pass, per pass:
 * offset mask, pool base, row count}.  Buffers are sized by ls3d_tile_plan_bytes and owned by the caller (16-byte aligned);
 * pool_counter is one device int of scratch. */
int ls3d_tile_plan_bytes(int32_t koff, int32_t m_out, int64_t* hdr_bytes, int64_t* local_bytes, int64_t* pool_bytes);
int ls3d_tile_plan_build(const int32_t* nbr, int32_t koff, int32_t m_out, int32_t* hdr, void* local, int32_t* pool,
                         int32_t* pool_counter, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hard voxelization of a batch of frames (bit-exact with the reference's numba voxelizer).
 * replaces: points_to_voxel / _points_to_voxel_reverse_kernel
 *           (det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184) called by
 *           VoxelGenerator.generate (det3d/core/input/voxel_generator.py:19-30) from SegVoxelization
 *           (det3d/datasets/pipelines/segpreprocess.py:148-177) + collate_kitti's batch-index padding
 *           (det3d/torchie/parallel/collate.py:141-150); API twin of hard_voxelize
 *           (det3d/ops/voxel/src/voxelization.cpp:6-11, voxelization_cpu.cpp:105-142).
 *   points         [n_points, n_feat] fp32 (x, y, z, ...), frames concatenated
 *   frame_off_host [n_frames + 1] HOST ints, first point of each frame
 *   voxel_size[3], pc_range[6] HOST floats (x, y, z order)
 *   voxels   [>= n_points, max_points, n_feat] fp32 (rows past total_voxels untouched)
 *   coords   [>= n_points, 4] int32 (frame, z, y, x);  num_points [>= n_points] int32
 *   num_voxels [n_frames], total_voxels [1] int32 (device);  point_voxel [n_points] or NULL
 * ------------------------------------------------------------------------------------------------ */
int ls3d_voxelize_workspace_bytes(int64_t n_points, int32_t max_points, int32_t n_frames, int64_t* bytes);
int ls3d_voxelize(const float* points, int32_t n_points, int32_t n_feat, const int32_t* frame_off_host,
                  int32_t n_frames, const float* voxel_size, const float* pc_range, int32_t max_points,
                  int32_t max_voxels, void* workspace, int64_t workspace_bytes, float* voxels, int32_t* coords,
                  int32_t* num_points, int32_t* num_voxels, int32_t* total_voxels, int32_t* point_voxel,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel feature encoders.
 * replaces: MeanVoxelFeatureExtractor / ImprovedMeanVoxelFeatureExtractor / the descriptor, attention
 *           core and slot max of TransformerVoxelFeatureExtractor
 *           (det3d/models/readers/voxel_encoder.py:51-58,74-124,149-157,202-270).
 *   mode 0: out[m, F] mean;  1: out[m, F+8] 13-d style descriptor;  2: out[m*P, 2F+8] token inputs
 *   round_out (here and below): 1 = store tf32-rounded values (result feeds a gather-GEMM), 0 = exact fp32
 * ------------------------------------------------------------------------------------------------ */
int ls3d_vfe_descriptor(const float* voxels, const int32_t* num_points, int32_t m, int32_t P, int32_t F,
                        int32_t mode, float* out, int32_t ld_out, int32_t round_out, void* stream);
int ls3d_vfe_token_attn(const float* qkv, int32_t ld_qkv, int32_t m, int32_t P, int32_t n_head, int32_t d_head,
                        float* out, int32_t ld_out, int32_t round_out, void* stream);
int ls3d_vfe_token_max(const float* x, int32_t ld_x, int32_t m, int32_t P, int32_t E, float* out, int32_t ld_out,
                       int32_t round_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Rulebooks (spconv "indice pairs") via an occupancy bitmap {bits, rank prefix} per 32 cells.
 * replaces: spconv get_indice_pairs behind SubMConv3d / SparseConv3d / SparseInverseConv3d
 *           (det3d/models/backbones/scn_unet.py:15-20,39-46,89-160).
 *   coords rows are int32 (frame, z, y, x).  nbr tables are [K][m] int32, K = kz*ky*kx,
 *   offset index k = (kz*KY + ky)*KX + kx, value = input row or -1.
 *   ls3d_grid_build         : bitmap of `coords`; perm (optional) = rank -> row for unsorted rows
 *   ls3d_grid_build_strided : bitmap of the strided conv's output sites; *total_out = #sites
 *   ls3d_grid_enumerate     : coords of the active cells in ascending linear order
 *   ls3d_rulebook_gather    : nbr[k][j] = input row at out_j*stride - pad + k  (SubM: stride 1, pad K/2)
 *   ls3d_rulebook_scatter   : inverse conv: nbr[k][i] = coarse row o with o*stride - pad + k == fine_i
 * ------------------------------------------------------------------------------------------------ */
int ls3d_grid_bytes(int32_t B, int32_t D, int32_t H, int32_t W, int64_t* words_bytes, int64_t* scratch_bytes);
int ls3d_grid_build(const int32_t* coords, int32_t m, int32_t B, int32_t D, int32_t H, int32_t W, void* words,
                    int32_t* perm, void* scratch, int32_t* total_out, void* stream);
int ls3d_grid_build_strided(const int32_t* in_coords, int32_t m_in, int32_t B, const int32_t* ksize,
                            const int32_t* stride, const int32_t* pad, int32_t oD, int32_t oH, int32_t oW,
                            void* out_words, void* scratch, int32_t* total_out, void* stream);
int ls3d_grid_enumerate(const void* words, int32_t B, int32_t D, int32_t H, int32_t W, int32_t* coords,
                        void* stream);
int ls3d_rulebook_gather(const void* in_words, const int32_t* in_perm, int32_t B, int32_t D, int32_t H, int32_t W,
                         const int32_t* out_coords, int32_t m_out, const int32_t* ksize, const int32_t* stride,
                         const int32_t* pad, int32_t* nbr, void* stream);
int ls3d_rulebook_scatter(const void* out_words, int32_t B, int32_t oD, int32_t oH, int32_t oW,
                          const int32_t* in_coords, int32_t m_in, const int32_t* ksize, const int32_t* stride,
                          const int32_t* pad, int32_t* nbr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Devoxelization: exact 3-NN of raw points among voxel centres + inverse-distance interpolation.
 * replaces: three_nn_wrapper / three_interpolate_wrapper (det3d/ops/pointnet2_batch/src/pointnet2_api.cpp:10-24,
 *           interpolate_gpu.cu:16-59,84-104) as used by three_interpolate_wrap
 *           (det3d/models/point_heads/point_utils.py:8-52).
 *   points [n, ld_p] fp32 rows (frame, x, y, z, ...); words/perm = level-1 bitmap of the voxel grid;
 *   idx are GLOBAL voxel rows (reference: per-frame rows; subtract voxel_off[frame] to compare);
 *   dist2 = squared distances (reference three_nn returns sqrt of these).
 *   workspace: todo int32[2 n] + todo_count int32[2] (work lists of the points handed from the per-thread box search to the
 *   warp-cooperative box search and from there to the exact brute-force pass).
 * ------------------------------------------------------------------------------------------------ */
/* Reference-signature twin: any point sets, b batches of n unknown / m known points [b, n|m, 3]; dist2 [b, n, 3] squared
 * distances, idx [b, n, 3] per-batch rows - exactly what three_nn_wrapper_fast(b, n, m, unknown, known, dist2, idx) fills
 * (interpolate.cpp:17-30; sequential strict-'<' scan, slots never filled keep (inf, 0)).  The lattice version below is the
 * fast path of the forward; this one is the drop-in for other callers of the pointnet2 extension. */
int ls3d_three_nn(int32_t b, int32_t n, int32_t m, const float* unknown, const float* known, float* dist2, int32_t* idx,
                  void* stream);
int ls3d_three_nn_grid(const float* points, int32_t ld_p, int32_t n, const void* words, const int32_t* perm, int32_t B,
                       int32_t D, int32_t H, int32_t W, const float* voxel_size_xyz, const float* range_min_xyz,
                       const int32_t* point_off, const int32_t* voxel_off, const int32_t* voxel_coords, int32_t* todo,
                       int32_t* todo_count, float* dist2, int32_t* idx, void* stream);
int ls3d_three_interpolate(const float* feat, int32_t ld_f, int32_t C, const float* dist2, const int32_t* idx, int32_t n,
                           float* out, int32_t ld_out, int32_t round_out, void* stream);
/* Segment table of a batch-sorted tensor: off[b] = first row whose batch column (fp32 when is_float, else int32; `stride`
 * elements between rows) is >= b, off[n_frames] = n.  What the reference obtains with per-frame boolean masks
 * (det3d/models/point_heads/point_utils.py:19-21, context_module.py:38-47). */
int ls3d_frame_offsets(const void* batch_col, int32_t is_float, int64_t stride, int32_t n, int32_t n_frames, int32_t* off,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera feature sampling.
 * replaces: PointSegMSeg3DHead.get_points_image_feature (F.grid_sample 3-D, bilinear, zeros,
 *           align_corners=True; det3d/models/point_heads/point_seg_mseg3d_head.py:200-236).
 *   feat_nhwc [n_frames, ncam, H, W, C] fp32 (feat_fp16 = 0) or fp16 (feat_fp16 = 1, fp32 arithmetic);
 *   points_cuv [n, 4] = (valid, cam, v, u) in [-1, 1]; out [n, C] fp32
 * ------------------------------------------------------------------------------------------------ */
int ls3d_sample_image_features(const void* feat_nhwc, int32_t feat_fp16, int32_t n_frames, int32_t ncam, int32_t H, int32_t W,
                               int32_t C, const float* points_cuv, int32_t n, const int32_t* point_off, float* out,
                               int32_t ld_out, int32_t round_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SF-Phase: point -> class-token cross attention core.
 * replaces: SparsePointCorssAttention.forward, attention part (det3d/models/point_heads/context_module.py:320-376);
 *           the q projection and the output projection around it are ls3d_gather_gemm launches.
 *   q [n, ld_q] fp32 (n_head * d_head columns used), k / v [n_frames][n_head][n_tok][d_head] from ls3d_class_tokens,
 *   frame_off[n_frames] int32 first row of each frame; out [n, ld_out]; n_head in {1,2,4,8}, d_head in {16,24,32}
 * ------------------------------------------------------------------------------------------------ */
int ls3d_token_attention(const float* q, int32_t ld_q, int32_t n, const float* k, const float* v, const int32_t* frame_off,
                         int32_t n_frames, int32_t n_tok, int32_t n_head, int32_t d_head, float scale, float* out,
                         int32_t ld_out, int32_t round_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SF-Phase: the whole POINT stream of the TransformerDecoder in one persistent launch (csrc/sffm_decoder.cu).
 * replaces: TransformerDecoder.forward layer loop + norm_tgt (det3d/models/point_heads/context_module.py:147-171),
 *           TransformerDecoderLayer.forward_post, point side (context_module.py:211-250),
 *           SparsePointCorssAttention.forward (context_module.py:320-376) - i.e. per layer q_proj, the class-token cross
 *           attention, out_proj + residual + norm2, linear1 + ReLU, linear2 + residual + norm3.
 *   tgt_in [n, ld_in] fp32 = input_proj_point output; out [n, ld_out] fp32 (after norm_tgt when final_norm != 0)
 *   w   : per layer the bf16x3 weight images (the gather-GEMM's packed layout, see ls3d_gemm_args.w, precise = 2) of
 *         q_proj (3 chunks), out_proj (3), linear1 (3 chunks of 192 rows x 128 B), linear2 (6), concatenated in that order;
 *         ls3d_sffm_decoder_weight_bytes gives the total size
 *   vec : per layer 864 floats [q bias 96 | out bias 96 | norm2 gamma 96, beta 96 | linear1 bias 192 | linear2 bias 96 |
 *         norm3 gamma 96, beta 96], then norm_tgt gamma 96, beta 96
 *   k, v [n_layer][n_frames][n_head][n_tok][d_model / n_head] from ls3d_class_tokens; frame_off[n_frames] first row per frame
 *   supported: d_model 96, d_ffn 192, n_head 4, n_layer <= 8, n_tok <= 64 (the shipped MSeg3D configs); else LS3D_ERR_ARG
 * ------------------------------------------------------------------------------------------------ */
int ls3d_sffm_decoder_weight_bytes(int32_t n_layer, int64_t* w_bytes, int64_t* vec_floats);
int ls3d_sffm_decoder(const float* tgt_in, int32_t ld_in, int32_t n, const void* w, const float* vec, const float* k,
                      const float* v, const int32_t* frame_off, int32_t n_frames, int32_t n_tok, int32_t n_layer,
                      int32_t n_head, int32_t d_model, int32_t d_ffn, int32_t final_norm, float attn_scale, float ln_eps,
                      float* out, int32_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Point -> camera projection (on the CPU, in the data loader, in the reference).
 * replaces: nuScenes branch of LoadPointCloudFromFile (det3d/datasets/pipelines/loading.py:373-416, view_points :67-103)
 *           + rescale / normalisation of SegImagePreprocess (det3d/datasets/pipelines/segpreprocess.py:544-565,654-671).
 *   points [n, ld_p] fp32 with x,y,z at columns xyz_off..xyz_off+2; cam_from_lidar [ncam][4][4], intrinsics [ncam][3][3]
 *   HOST doubles (row-major); points_cuv [n, 4] = (valid, cam, v, u), cam/v/u normalised to [-1, 1]
 * ------------------------------------------------------------------------------------------------ */
int ls3d_project_points(const float* points, int32_t ld_p, int32_t xyz_off, int32_t n, const double* cam_from_lidar,
                        const double* intrinsics, int32_t ncam, int32_t img_h, int32_t img_w, int32_t net_h, int32_t net_w,
                        float* points_cuv, void* stream);
/* Same with the loader's two-stage chain lidar -> global -> camera (info["ref_to_global"] [4][4], then
 * info["cams_from_global"][cam] [ncam][4][4]; loading.py:386-395), each stage rounded to fp64 like the numpy original. */
int ls3d_project_points_global(const float* points, int32_t ld_p, int32_t xyz_off, int32_t n, const double* ref_to_global,
                               const double* cams_from_global, const double* intrinsics, int32_t ncam, int32_t img_h,
                               int32_t img_w, int32_t net_h, int32_t net_w, float* points_cuv, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera stem: multi-resolution branch fusion, out = act(bias + sum_k bilinear_resize(term_k)) on channels-last fp32 maps.
 * replaces: the per-term resize / add / ReLU loops of HRModule.forward (det3d/models/img_backbones/hrnet.py:205-226) and
 *           the resize_concat of the decode head (det3d/models/img_heads/decode_head.py:151-160) once its 1x1 ConvModule
 *           has been applied per branch (fcn_mseg3d_head.py:150-163).
 *   terms    HOST array of n_terms (<= 4) device pointers, term k = [n_img, term_h[k], term_w[k], C] NHWC; a term with
 *            term_h == H and term_w == W is added as is, a coarser one is resized with align_corners = False
 *   bias     [C] fp32 on the device or NULL: added once per output element (the summed folded-BatchNorm shifts of the
 *            bias-free convolutions behind the terms; a per-channel constant commutes with the resize)
 *   out      [n_img, H, W, C]; C a multiple of 4; relu != 0 applies max(., 0) to the sum
 * ------------------------------------------------------------------------------------------------ */
int ls3d_upsample_sum(const float* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                      int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, float* out, void* stream);
/* same on fp16 maps (fp32 arithmetic); C a multiple of 8 */
int ls3d_upsample_sum_f16(const void* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                          int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, void* out, void* stream);
int ls3d_upsample_sum_dual(const float* const* terms, const int32_t* term_h, const int32_t* term_w, int32_t n_terms,
                           int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t relu, const float* bias, float* out,
                           void* out16, void* stream);   /* fp32 result + its fp16 operand copy in one pass */
/* fp16 (round-to-nearest-even) copy of n fp32 values, n a multiple of 4: the tensor-core operand copy of an fp32 map */
int ls3d_cast_f16(const float* in, void* out, int64_t n, void* stream);
/* fp32 [n_pixels][3] (channels-last network input) -> fp16 [n_pixels][8] with channels 3..7 zero: operand copy of the image for
 * the own stem convolution (hrnet.py:658-666), whose tensor-map copies need 16-byte pixel rows */
int ls3d_pad3_f16(const float* in, int64_t n_pixels, void* out, void* stream);
/* fp32 copy of n fp16 values (n a multiple of 4) */
int ls3d_cast_f32(const void* in, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera input preparation: uint8 [n_pixels][3] (HWC images, any batch of them back to back) -> (x / 255 - mean[c]) / std[c]
 * in IEEE fp32 (mean / std as fp32: within 1 ulp per operation of the reference's numpy, whose in-place ops promote its
 * python-float mean / std to fp64), stored fp32 or fp16 in the same pixel-major order (= channels-last maps).
 * replaces: image_input_transform (det3d/datasets/pipelines/img_transforms.py:18-29, segpreprocess.py:621-628) + the HWC->CHW
 *           transpose (segpreprocess.py:637); mean3 / std3 are HOST pointers to 3 floats (cam_attributes[cam]["mean"/"std"]).
 *   in and out 16-byte aligned.
 * ------------------------------------------------------------------------------------------------ */
int ls3d_normalize_images_u8(const uint8_t* in, int64_t n_pixels, const float* mean3, const float* std3, void* out,
                             int32_t out_fp16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera input: cv2.resize (uint8, INTER_LINEAR, bit-exact with OpenCV's fixed-point algorithm) fused with the normalisation.
 * replaces: cv2.resize in image_and_points_cp_and_label_resize (det3d/datasets/pipelines/img_transforms.py:78-99, called
 *           per camera by SegImagePreprocess.__call__, segpreprocess.py:544-565) followed by image_input_transform
 *           (img_transforms.py:18-29; segpreprocess.py:621-628) and the HWC -> CHW transpose (:637).
 *   in  [n_img, in_h, in_w, 3] uint8 (raw decoded camera images, e.g. 900 x 1600)
 *   out [n_img, out_h, out_w, 3]: out_kind 0 = normalised fp32, 1 = normalised fp16 (channels-last stem input),
 *                                 2 = resized uint8 only (mean3 / std3 ignored, may be NULL)
 *   mean3 / std3: HOST pointers to 3 floats.
 * ------------------------------------------------------------------------------------------------ */
int ls3d_resize_images_u8(const uint8_t* in, int32_t n_img, int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w,
                          const float* mean3, const float* std3, void* out, int32_t out_kind, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Camera stem: fused 3x3 / stride 1 / pad 1 convolution + bias (+ residual) (+ ReLU) on channels-last fp16 maps,
 * tcgen05 tensor cores, fp32 accumulation.
 * replaces: conv3x3 -> BatchNorm (folded) -> [+ identity] -> ReLU of the HRNet BasicBlocks
 *           (det3d/models/img_backbones/resnet_mmcv.py:20-100 as used by hrnet.py:78-226), cuDNN in the reference.
 *   in  [n_img, H, W, cin] fp16, res / out [n_img, H, W, cout] fp16 (res may be NULL); cin, cout multiples of 8 (zero-padded
 *   channels), all base pointers 16-byte aligned; bias [cout] fp32 or NULL.
 *   w_packed: ls3d_conv3x3_f16_packed_bytes(cin, cout) bytes written by ls3d_conv3x3_f16_pack from the BatchNorm-folded fp32
 *   weights [cout][cin][3][3] (the kernel's K order pairs (tap, 8-channel chunk) entries two per MMA; opaque to the caller).
 *   ls3d_conv3x3_f16_smem_bytes: shared memory the launch needs (weights stay resident); > 227 KB = not supported.
 * ------------------------------------------------------------------------------------------------ */
int ls3d_conv3x3_f16_smem_bytes(int32_t cin, int32_t cout, int64_t* bytes);
int ls3d_conv3x3_f16_packed_bytes(int32_t cin, int32_t cout, int64_t* bytes);
int ls3d_conv3x3_f16_pack(const float* w_oihw, int32_t cin, int32_t cout, void* packed, void* stream);
int ls3d_conv3x3_f16(const void* in, const void* w_packed, const float* bias, const void* res, void* out, int32_t n_img,
                     int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t relu, void* stream);
/* the same kernel for ksize = 3 (above) or ksize = 1 (1x1 / stride 1 / pad 0: only the centre of the halo enters the K list;
 * weights [cout][cin][1][1]) - the 1x1 convolutions of the HRNet fuse layers / Bottlenecks (hrnet.py:156-204,
 * resnet_mmcv.py:103-225) and of the FCN decode head (fcn_mseg3d_head.py:150-163), cuBLAS/cuDNN in the reference */
int ls3d_conv_f16_smem_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes);
int ls3d_conv_f16_packed_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes);
int ls3d_conv_f16_pack(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, void* packed, void* stream);
int ls3d_conv_f16(const void* in, const void* w_packed, const float* bias, const void* res, void* out, int32_t n_img, int32_t H,
                  int32_t W, int32_t cin, int32_t cout, int32_t ksize, int32_t relu, void* stream);
/* fp32 feature maps with fp16 tensor-core operands ("fp32 residual stream"): in16 = fp16 operand copy of the input map,
 * res32 (may be NULL) / out32 = fp32 maps, out16 = fp16 operand copy of the result for the next convolution.  The
 * multiplications see the 11-bit significand a TF32 tensor-core convolution (the reference's stock cuDNN path) sees;
 * accumulation, bias, residual, ReLU and the stored maps are fp32.
 *   out32 == NULL: operand-only result (only out16 is written, res32 must be NULL): a map that only feeds the next convolution
 *                  (conv1 of a BasicBlock).
 *   w_split != 0 : w_packed holds SPLIT weights [W_hi ; W_lo] (ls3d_conv_f16_pack_split: fp16(w) and fp16(w - fp16(w)) stacked
 *                  along N, one MMA per K slice over 2 n_pad columns, the epilogue adds the halves): the fp32 weights enter
 *                  exactly (2^-22), so the only rounding left is the RNE fp16 operand copy of the activations - an unbiased,
 *                  spatially incoherent error (measured on the CPU oracle, scripts/sim_operand_rounding.py: 3.4e-5 on the
 *                  logits against 3.6e-4 when the weights are rounded too).  Needs 2 n_pad <= 256 and twice the resident weight
 *                  bytes (ls3d_conv_f16_split_supported). */
int ls3d_conv_f16_dual(const void* in16, const void* w_packed, const float* bias, const float* res32, float* out32, void* out16,
                       int32_t n_img, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t ksize, int32_t relu,
                       int32_t w_split, void* stream);
int ls3d_conv_f16_dual_smem_bytes(int32_t cin, int32_t cout, int32_t ksize, int64_t* bytes);
int ls3d_conv_f16_split_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t dual, int32_t* supported);
int ls3d_conv_f16_pack_split(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, void* packed, void* stream);

/* General form of the same kernel: every convolution of the camera branch (HRNet stem / Bottlenecks / transitions / fuse layers,
 * FCN head; reference hrnet.py:156-226,658-693, resnet_mmcv.py:20-225, fcn_mseg3d_head.py:150-163) is one or a few launches.
 *   stride 2 (3x3, pad 1): the halo holds the four (row parity, column parity) phase planes of the input, each read through
 *       its own tensor map over the same memory with doubled pixel strides; tap (dy, dx) = a shifted view of one plane.
 *   channel slices: the launch convolves input channels [in_c_off, in_c_off + cin) of a tensor with in_c_total channels into
 *       output channels [out_c_off, out_c_off + cout) of tensors with out_c_total channels.  Shapes whose weights do not fit
 *       shared memory run as several launches over input slices that ACCUMULATE through the fp32 residual input
 *       (res32 == out32 is allowed: each tile reads its residual before it is written); > 128 output channels with split
 *       weights run as output slices.
 *   H_in, W_in: input map size; the output is H_in x W_in (stride 1) or ceil(H_in / 2) x ceil(W_in / 2) (stride 2).
 *   w_packed: ls3d_conv_f16_pack_ex of the [cout][cin][k][k] weight slice with the same (ksize, stride, w_split). */
typedef struct ls3d_conv_args {
  const void* in16;       /* fp16 operand copy of the input map, channels-last [n_img, H_in, W_in, in_c_total] */
  const void* w_packed;
  const float* bias;      /* [cout] fp32 or NULL */
  const float* res32;     /* fp32 [n_img, H_out, W_out, out_c_total] or NULL */
  float* out32;           /* fp32 output map or NULL (operand-only launch: out16 alone, no residual) */
  void* out16;            /* fp16 operand copy of the output map */
  int32_t in_c_total, in_c_off, cin;
  int32_t out_c_total, out_c_off, cout;
  int32_t n_img, H_in, W_in, ksize, stride, relu, w_split;
  const void* res16;      /* fp16 residual [n_img, H_out, W_out, out_c_total] for operand-only launches (out32 == NULL): fp16 maps */
} ls3d_conv_args;
int ls3d_conv_f16_ex(const ls3d_conv_args* args, void* stream);
/* The launches of a sliced convolution as PASSES of one launch (a CTA owns the same tiles in every pass, so accumulation through
 * the fp32 output map stays inside the CTA): args->cin / cout = the uniform slice sizes (slices may overlap: the overlapping
 * weights of the later slice are zero / the overlapping outputs are recomputed), args->bias = bias of the WHOLE output tensor.
 *   flags: 1 add bias[out_c_off ...], 2 add the external residual (res32 / res16), 4 add the output map itself (what the
 *          previous pass left: needs out32), 8 ReLU.  At most 24 passes. */
typedef struct ls3d_conv_pass {
  const void* w_packed;
  int32_t in_c_off, out_c_off, flags;
} ls3d_conv_pass;
int ls3d_conv_f16_multi(const ls3d_conv_args* args, const ls3d_conv_pass* passes, int32_t n_pass, void* stream);
/* Streamed-weight variant for the weight-heavy, pixel-light 3x3 convolutions, stride 1 or 2 (HRNet's 72- and 144-channel branches
 * and the stride-2 fuse convolutions that feed them):
 * every pass = one OUTPUT channel slice over all input channels (in_c_off = 0, flags without 4); a work item is (group of up to
 * 4 output tiles, slice): the group's halos stay staged, the slice's packed weights (ls3d_conv_f16_pack_ex image, unchanged)
 * stream once per item through a shared-memory ring and every MMA runs at the slice's full N.  args->cout = slice size
 * (<= 128), at most 8 slices.  Its epilogue reads the residual and writes the maps straight from / to global memory (no staged
 * output tile), which also makes it the kernel of the wide 1x1 convolutions (Bottleneck conv3, 64 -> 256: 128-channel slices).  ls3d_conv_f16_kb_supported: is there a shared / tensor-memory configuration for this shape? */
int ls3d_conv_f16_kb(const ls3d_conv_args* args, const ls3d_conv_pass* passes, int32_t n_slices, void* stream);
int ls3d_conv_f16_kb_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t dual, int32_t split,
                               int64_t n_out_pixels, int32_t* supported);
int ls3d_conv_f16_ex_supported(int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t dual, int32_t split,
                               int32_t* supported);
int ls3d_conv_f16_pack_ex(const float* w_oihw, int32_t cin, int32_t cout, int32_t ksize, int32_t stride, int32_t split,
                          void* packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SF-Phase: class embedding aggregation and class-token memory path.
 * replaces: LiDARSemanticFeatureAggregationModule (det3d/models/point_heads/context_module.py:25-53),
 *           CameraSemanticFeatureAggregationModule (det3d/models/img_heads/fcn_mseg3d_head.py:24-51),
 *           memory side of SemanticFeatureFusionModule / TransformerDecoderLayer.forward_post
 *           (context_module.py:101-109,211-227,337-339).
 *   logits [rows][ld_l], feats [rows][ld_f]: both fp32 (in_fp16 = 0) or both fp16 (in_fp16 = 1; fp32 arithmetic)
 *   emb  [n_frames][ncls][C];  K, V [n_layer][n_frames][n_head][2*ncls][d_model/n_head]
 * ------------------------------------------------------------------------------------------------ */
int ls3d_class_embed_workspace_bytes(int32_t n_frames, int32_t max_rows_per_frame, int32_t ncls, int32_t C,
                                     int64_t* bytes);
int ls3d_class_embed(const void* logits, int32_t ld_l, int32_t ncls, const void* feats, int32_t ld_f, int32_t C,
                     int32_t in_fp16, const int32_t* seg_off, int32_t n_frames, int32_t max_rows_per_frame, void* workspace,
                     float* emb, void* stream);
int ls3d_class_tokens(const float* emb1, int32_t C1, const float* emb2, int32_t C2, int32_t ncls, int32_t n_frames,
                      const float* params, int32_t n_layer, int32_t n_head, int32_t d_model, float* K, float* V,
                      float* mem_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS3D_H_ */
