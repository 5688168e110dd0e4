// ls3d_gather_gemm: entry point and argument checks of the gather-GEMM family (sm_100a).
//
// One contract serves every dense contraction of the MSeg3D / SDSeg3D forward path:
//   * sparse 3-D convolution (SubM / strided / inverse) in output-stationary form
//       out[j,:] = epi( sum_k  in[nbr[k][j], :] . W[k] )          (replaces spconv's per-offset
//       gather -> cuBLAS mm -> scatter-add triple, reference call sites
//       det3d/models/backbones/scn_unet.py:15-20,39-46,89-160)
//   * every Linear(+BN)(+ReLU)(+residual)(+LayerNorm) of the point heads / TransVFE (koff = 1, nbr = identity), reference
//       det3d/models/point_heads/*.py, det3d/models/readers/voxel_encoder.py:128-270
//   * the SF-Phase class-token cross attention as an epilogue of the q projection
//       (reference det3d/models/point_heads/context_module.py:320-376; the shipped head uses ls3d_sffm_decoder instead).
// Engines (error-compensated bf16x3 products, fp32 accumulation in tensor memory, epilogues of gemm_epilogue.cuh):
//   gather_gemm_once.cu   - launches that carry a tile plan (every sparse convolution): distinct rows staged once per tile
//   gather_gemm_bf16x3.cu - per-pair gathers: the dense Linears, and sparse launches whose plan has no shared-memory fit
// The round-1 TF32 / 3xTF32 engines that lived in this file were retired once the bf16x3 engines served every launch.
#include "gemm_epilogue.cuh"

int ls3d_gather_gemm_bf16x3_launch(const ls3d_gemm_args* a, int num_sms, void* stream);   // gather_gemm_bf16x3.cu
int ls3d_gather_gemm_once_launch(const ls3d_gemm_args* a, int num_sms, void* stream);     // gather_gemm_once.cu
int ls3d_gather_gemm_once_fits(const ls3d_gemm_args* a);

extern "C" int ls3d_gather_gemm(const ls3d_gemm_args* a, void* stream) {
  using namespace ls3d;
  if (!a || !a->in0 || !a->w || !a->out) return LS3D_ERR_ARG;
  if (a->m_out <= 0) return LS3D_OK;
  if (a->precise != 2) return LS3D_ERR_ARG;                       // the only engine family: bf16x3 (see include/ls3d.h)
  if (a->koff < 1 || a->koff > MAX_KOFF) return LS3D_ERR_ARG;
  if (!a->nbr && a->koff != 1) return LS3D_ERR_ARG;
  if (a->n_pad % 16 || a->n_pad < 16 || a->n_pad > 256 || a->cout > a->n_pad) return LS3D_ERR_ARG;
  if (a->cin_pad % 16 || a->cin_pad < a->c0 + a->c1) return LS3D_ERR_ARG;
  if ((a->c0 & 3) || (a->c1 & 3) || (a->ld0 & 3) || (a->c1 && (a->ld1 & 3))) return LS3D_ERR_ARG;
  if (a->epi == LS3D_EPI_ATTN) {
    if (!a->attn_k || !a->attn_v || !a->frame_off || a->n_tok > MAX_TOK ||
        a->n_head * DHEAD != a->cout || (a->ld_out & 3))
      return LS3D_ERR_ARG;
  }
  if (a->n_ln < 0 || a->n_ln > 2) return LS3D_ERR_ARG;
  if (a->red0 && (!a->red1 || (a->red_c & 3) || (a->ld_red0 & 3) || (a->ld_red1 & 3) || a->cout != a->red_c ||
                  a->epi != LS3D_EPI_LINEAR))
    return LS3D_ERR_ARG;
  const int num_sms = ls3d_num_sms();
  if (a->nbr && a->plan_hdr && ls3d_gather_gemm_once_fits(a)) return ls3d_gather_gemm_once_launch(a, num_sms, stream);
  return ls3d_gather_gemm_bf16x3_launch(a, num_sms, stream);
}
