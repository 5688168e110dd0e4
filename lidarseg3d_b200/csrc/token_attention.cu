// SF-Phase point -> class-token cross attention: out[r, h] = softmax(q[r, h] . K_f(r)[h]^T * scale) V_f(r)[h].
//
// Replaces the attention core of SparsePointCorssAttention.forward (reference det3d/models/point_heads/context_module.py:
// 320-376: per frame, q [n_pts, H, dh] against that frame's n_tok class tokens, softmax over the tokens, weighted value sum).
// The q projection before it and the output projection after it are ls3d_gather_gemm launches; K and V are the per-layer
// token projections of ls3d_class_tokens ([frame][head][token][dh], a few KB per frame).
//
// Layout: one thread per (point, head).  A block covers `rows_per_block` consecutive points for all heads (head = warp-
// uniform), stages the K/V of the frame of its first point in shared memory (broadcast LDS.128 reads) and streams the
// tokens in chunks of 8 with a running max / denominator, so the register footprint does not depend on n_tok.  A point
// whose frame differs from the staged one (only in a block straddling a frame boundary) reads K/V from global memory.
// Work per point: 4*n_tok*dh FMAs against 2*H*dh*4 bytes in+out -> compute/issue bound on the CUDA cores, not HBM.
#include "common.cuh"
#include "../../include/ls3d.h"

namespace ls3d {

constexpr int TA_THREADS = 256;
constexpr int TA_CHUNK = 8;

template <int DH>
__device__ __forceinline__ void token_attention_row(const float* __restrict__ qrow, const float* kh, const float* vh, int L,
                                                    float scale, int round_out, float* __restrict__ orow) {
  float q[DH];
#pragma unroll
  for (int d4 = 0; d4 < DH / 4; ++d4) {
    const float4 t = ldg_f4(qrow + d4 * 4);
    q[d4 * 4] = t.x; q[d4 * 4 + 1] = t.y; q[d4 * 4 + 2] = t.z; q[d4 * 4 + 3] = t.w;
  }
  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float mx = -INFINITY, den = 0.f;
  for (int l0 = 0; l0 < L; l0 += TA_CHUNK) {
    float s[TA_CHUNK];
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < TA_CHUNK; ++j) {
      float a = -INFINITY;
      if (l0 + j < L) {
        const float4* kr = reinterpret_cast<const float4*>(kh + (size_t)(l0 + j) * DH);
        a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 kv = kr[d4];
          a = fmaf(q[d4 * 4 + 0], kv.x, a);
          a = fmaf(q[d4 * 4 + 1], kv.y, a);
          a = fmaf(q[d4 * 4 + 2], kv.z, a);
          a = fmaf(q[d4 * 4 + 3], kv.w, a);
        }
        a *= scale;
      }
      s[j] = a;
      cm = fmaxf(cm, a);
    }
    const float mn = fmaxf(mx, cm);
    const float corr = __expf(mx - mn);            // 0 on the first chunk (mx = -inf)
    den *= corr;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < TA_CHUNK; ++j) {
      if (l0 + j < L) {
        const float e = __expf(s[j] - mn);
        den += e;
        const float4* vr = reinterpret_cast<const float4*>(vh + (size_t)(l0 + j) * DH);
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 vv = vr[d4];
          o[d4 * 4 + 0] = fmaf(e, vv.x, o[d4 * 4 + 0]);
          o[d4 * 4 + 1] = fmaf(e, vv.y, o[d4 * 4 + 1]);
          o[d4 * 4 + 2] = fmaf(e, vv.z, o[d4 * 4 + 2]);
          o[d4 * 4 + 3] = fmaf(e, vv.w, o[d4 * 4 + 3]);
        }
      }
    }
    mx = mn;
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int d4 = 0; d4 < DH / 4; ++d4) {
    float4 t = make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
    if (round_out) t = make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
    *reinterpret_cast<float4*>(orow + d4 * 4) = t;
  }
}

template <int DH>
__global__ void __launch_bounds__(TA_THREADS) token_attention_kernel(const float* __restrict__ q, int ld_q, int n,
                                                                     const float* __restrict__ k, const float* __restrict__ v,
                                                                     const int32_t* __restrict__ frame_off, int n_frames, int L,
                                                                     int H, float scale, float* __restrict__ out, int ld_out,
                                                                     int round_out) {
  extern __shared__ __align__(16) float ta_smem[];
  const int rpb = TA_THREADS / H;                       // points per block; a multiple of 32, so `head` is warp-uniform
  const int head = threadIdx.x / rpb, rl = threadIdx.x % rpb;
  const int row0 = blockIdx.x * rpb;
  auto frame_of = [&](int r) {
    int f = 0;
    for (int i = 1; i < n_frames; ++i)
      if (r >= frame_off[i]) f = i;
    return f;
  };
  const int f0 = frame_of(row0);
  const int per_frame = H * L * DH;                     // floats of K (and of V) per frame
  float* sk = ta_smem;
  float* sv = ta_smem + per_frame;
  {
    const float4* gk = reinterpret_cast<const float4*>(k + (size_t)f0 * per_frame);
    const float4* gv = reinterpret_cast<const float4*>(v + (size_t)f0 * per_frame);
    for (int i = threadIdx.x; i < per_frame / 4; i += TA_THREADS) {
      reinterpret_cast<float4*>(sk)[i] = __ldg(gk + i);
      reinterpret_cast<float4*>(sv)[i] = __ldg(gv + i);
    }
  }
  __syncthreads();
  const int r = row0 + rl;
  if (r >= n) return;
  const int f = frame_of(r);
  const float* qrow = q + (size_t)r * ld_q + head * DH;
  float* orow = out + (size_t)r * ld_out + head * DH;
  if (f == f0) {
    token_attention_row<DH>(qrow, sk + (size_t)head * L * DH, sv + (size_t)head * L * DH, L, scale, round_out, orow);
  } else {
    const size_t off = ((size_t)f * H + head) * L * DH;
    token_attention_row<DH>(qrow, k + off, v + off, L, scale, round_out, orow);
  }
}

}  // namespace ls3d

extern "C" int ls3d_token_attention(const float* q, int32_t ld_q, int32_t n, const float* k, const float* v,
                                    const int32_t* frame_off, int32_t n_frames, int32_t n_tok, int32_t n_head, int32_t d_head,
                                    float scale, float* out, int32_t ld_out, int32_t round_out, void* stream) {
  using namespace ls3d;
  if (n <= 0) return LS3D_OK;
  if (!q || !k || !v || !frame_off || !out || n_frames < 1 || n_tok < 1) return LS3D_ERR_ARG;
  if (n_head != 1 && n_head != 2 && n_head != 4 && n_head != 8) return LS3D_ERR_ARG;
  if ((ld_q & 3) || (ld_out & 3)) return LS3D_ERR_ARG;
  const size_t smem = (size_t)2 * n_head * n_tok * d_head * sizeof(float);
  if (smem > 200 * 1024) return LS3D_ERR_ARG;
  const int rpb = TA_THREADS / n_head;
  const int grid = ls3d_div_up(n, rpb);
  cudaStream_t st = (cudaStream_t)stream;
#define LS3D_TA_LAUNCH(DH)                                                                                              \
  do {                                                                                                                  \
    if (smem > 48 * 1024)                                                                                               \
      cudaFuncSetAttribute(token_attention_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    token_attention_kernel<DH><<<grid, TA_THREADS, smem, st>>>(q, ld_q, n, k, v, frame_off, n_frames, n_tok, n_head, scale, \
                                                               out, ld_out, round_out);                                \
  } while (0)
  if (d_head == 24) LS3D_TA_LAUNCH(24);
  else if (d_head == 16) LS3D_TA_LAUNCH(16);
  else if (d_head == 32) LS3D_TA_LAUNCH(32);
  else return LS3D_ERR_ARG;
#undef LS3D_TA_LAUNCH
  LS3D_LAUNCH_CHECK();
  return LS3D_OK;
}
