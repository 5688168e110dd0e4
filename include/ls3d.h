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
 * Gather-GEMM (tcgen05, TF32 in / fp32 accumulate) with fused epilogue.
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
  const float* w;     /* [koff][n_pad][cin_pad] (K-major), tf32-rounded, zero padded           */
  int32_t cin_pad;    /* multiple of 8, >= c0 + c1                                             */
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
  float* out;         /* [m_out, ld_out]                                                       */
  int32_t ld_out;
} ls3d_gemm_args;

int ls3d_gather_gemm(const ls3d_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LS3D_H_ */
